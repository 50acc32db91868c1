"""GPU parity of the fused env_shade kernel AT THE BASELINE SIZES: every case in test_gpu_envshade.py is
<= 24 x 24 pixels, which never fills the persistent grid (592 CTAs), never steals work across CTAs, never needs more than one
queue fill per pixel at scale and never writes a multi-GB ray record.  Here the product runs the full-size launch and the oracle
re-computes a subset of its pixels (tests/parity_check.py): integer records bit-exact, radiance and all five gradients <= 1e-4.

  * configs[1]  bob-like:  1 x 512^2, n_samples_x = 4, 11 264 triangles, random 256^2 probe
  * configs[2]  spot-like: 8 x 256^2, n_samples_x = 8, 7 168 triangles, HDR 256 x 512 probe with a 900x sun (light-gradient atomics
                hot spot) -- the 8 x 512^2 launch itself is checked by bench.py --verify on every bench run
  * configs[4]  1 M-triangle grid: 1 x 256^2, n_samples_x = 16 (512 sample slots = 4 queue fills per pixel, envshade.cu nsub)
"""
import numpy as np
import pytest
import torch

from common import oracle
from parity_check import env_shade_parity, select_pixels

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _gpu_case(dev, mesh, res, B, N, light, light_hw, seed=0, radius=3.0, tilt=-0.4, level=4, ks_mode="random", perm_rows=4096):
    import nvdiffrecmc_b200.optixutils as ou
    import nvdiffrecmc_b200.renderutils as ru
    from nvdiffrecmc_b200 import synth
    o = oracle()
    v, f = synth.scene_mesh(mesh, level=level, seed=5 + seed)
    vn = synth.vertex_normals(v, f)
    ctx = ou.OptiXContext()
    ou.optix_build_bvh(ctx, torch.tensor(v, device=dev), torch.tensor(f, device=dev), rebuild=1)
    gbs = []
    for b in range(B):
        mv = synth.orbit_view(0.7 * b + 0.3 * seed, radius=radius, tilt=tilt)
        campos, ro, rd = synth.primary_rays(mv, res)
        tid, tuv = ou.trace_closest(ctx, torch.tensor(ro.reshape(-1, 3), device=dev), torch.tensor(rd.reshape(-1, 3), device=dev))
        gbs.append(synth.assemble_gbuffer(v, f, vn, tid.cpu().numpy().reshape(res, res), tuv.cpu().numpy().reshape(res, res, 3), campos,
                                          seed=1 + b + 10 * seed, ks_mode=ks_mode))
    st = lambda k: torch.tensor(np.stack([g[k] for g in gbs]), device=dev)
    view = st("view_pos").reshape(B, 1, 1, 3)
    pos, mask = st("pos"), st("mask")
    nrm = ru.prepare_shading_normal(pos, view, None, st("smooth_nrm"), st("tangent"), st("geom_nrm"), two_sided_shading=True, opengl=True)
    nrm = (nrm * (mask[..., None] > 0)).contiguous()
    ro = (pos + nrm * 0.001).contiguous()
    base = synth.random_light(light_hw[0], seed=2 + seed) if light == "random" else synth.hdr_light(light_hw[0], light_hw[1], seed=7 + seed)
    pdf, rows, cols = o.update_pdf(base)
    t = lambda a_: torch.tensor(a_, device=dev)
    dev_in = dict(mask=mask, ro=ro, pos=pos, nrm=nrm, view=view, kd=st("kd"), ks=st("ks"), light=t(base), pdf=t(pdf), rows=t(rows), cols=t(cols))
    perms = t(synth.make_perms(N, seed=3 + seed, rows=perm_rows))
    return ctx, o.scene(v, f), dev_in, perms


def _assert_green(r, min_occluded=100):
    assert r["texel_mismatch"] == 0, r
    assert r["vis_mismatch"] == 0, r
    assert r["rays_occluded_checked"] >= min_occluded, r          # the selection really contains shadowed rays
    assert r["grad_outside_selection_max_abs"] == 0.0, r
    assert r["max_rel_l2"] <= TOL, r


def test_bob_like_512_n4(dev):
    N = 4
    ctx, scene, dev_in, perms = _gpu_case(dev, "bob-like", 512, 1, N, "random", (256, 256), seed=1)
    mask = dev_in["mask"].cpu().numpy()
    assert 0.15 < (mask > 0).mean() < 0.8
    sel = select_pixels(mask, crop=96, n_random=6000, seed=1)
    r = env_shade_parity(ctx, scene, dev_in, perms, N, sel, bsdf="pbr", seed=17)
    assert r["rays_checked"] > 200000
    _assert_green(r)


@pytest.mark.parametrize("bsdf,shadow_scale", [("pbr", 1.0), ("diffuse", 0.5)])
def test_spot_like_8_views_n8_hdr_probe(dev, bsdf, shadow_scale):
    N = 8
    ctx, scene, dev_in, perms = _gpu_case(dev, "blob+torus", 256, 8, N, "hdr", (256, 512), seed=2, ks_mode="metal" if bsdf == "pbr" else "random")
    mask = dev_in["mask"].cpu().numpy()
    sel = select_pixels(mask, crop=48, n_random=3000, seed=2)
    r = env_shade_parity(ctx, scene, dev_in, perms, N, sel, bsdf=bsdf, seed=23, shadow_scale=shadow_scale)
    assert r["rays_checked"] > 300000
    _assert_green(r)


def test_million_triangles_256_n16(dev):
    N = 16
    ctx, scene, dev_in, perms = _gpu_case(dev, "grid1m", 256, 1, N, "random", (256, 256), seed=3, radius=2.2, tilt=-0.9, perm_rows=1024)
    mask = dev_in["mask"].cpu().numpy()
    assert (mask > 0).mean() > 0.3
    sel = select_pixels(mask, crop=32, n_random=1500, seed=3)
    r = env_shade_parity(ctx, scene, dev_in, perms, N, sel, bsdf="pbr", seed=29)
    assert r["rays_checked"] > 500000
    _assert_green(r)


def test_batch_offset_global_view_index(dev):
    """The verify leg of bench.py at N > 1 GPUs relies on it: views [4, 8) of a global batch rendered with batch_offset=4."""
    N = 4
    ctx, scene, dev_in, perms = _gpu_case(dev, "blob+torus", 96, 2, N, "random", (64, 64), seed=4)
    sel = select_pixels(dev_in["mask"].cpu().numpy(), crop=24, n_random=300, seed=4)
    r = env_shade_parity(ctx, scene, dev_in, perms, N, sel, seed=31, batch_offset=4)
    _assert_green(r, min_occluded=10)

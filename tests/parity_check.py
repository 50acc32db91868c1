"""Parity of ONE product env_shade launch with the CPU oracle on a SUBSET of its pixels -- TEST INFRASTRUCTURE (imports oracle/).

Used by the `-m gpu` tests at the BASELINE sizes (tests/test_gpu_envshade_large.py) and by bench.py's `--verify` leg, which checks
the very launch configuration it times (8 x 512^2, n_samples_x = 8: persistent grid with work stealing, 1 M active pixels, the ray
record) instead of only toy sizes.

Every pixel of env_shade is independent given its (global view, y, x) index (kernel.cu:504 hashes the pixel index into the RNG
seed), so the oracle is run with the mask zeroed outside the selected pixels, view by view with batch_offset = global view index;
the product runs the FULL launch.  Integer records (env texel per sample, shadow-ray visibility bit) must agree bit for bit, radiance
and the five gradients within 1e-4 relative L2 (BASELINE.json north_star).  The env-map gradient is a sum over pixels: the product's
backward pass is driven with upstream gradients that are zero outside the selection, so its light gradient is the selection's.

Reference lines restated by the oracle: render/optixutils/c_src/envsampling/kernel.cu:463-542 (__raygen__rg), :403-461
(process_sample), :101-118 (shadow_test).
"""
import numpy as np
import torch

from common import oracle, rel_l2


def select_pixels(mask, crop=64, n_random=2048, seed=0):
    """bool [B,H,W]: view 0's centre crop x crop plus n_random covered pixels drawn from all views."""
    B, H, W = mask.shape
    sel = np.zeros((B, H, W), bool)
    c = min(crop, H, W)
    y0, x0 = (H - c) // 2, (W - c) // 2
    sel[0, y0:y0 + c, x0:x0 + c] = True
    cov = np.flatnonzero(mask.reshape(-1) > 0)
    if cov.size and n_random > 0:
        pick = np.random.default_rng(seed).choice(cov, size=min(n_random, cov.size), replace=False)
        sel.reshape(-1)[pick] = True
    return sel


def env_shade_parity(ctx, scene, dev_in, perms_dev, N, sel, bsdf="pbr", seed=0, shadow_scale=1.0, batch_offset=0, bench_fwd=None,
                     upstream_seed=1):
    """ctx: product OptiXContext with the BVH of `scene`'s mesh built; scene: oracle Scene of the same mesh;
    dev_in: dict of CUDA tensors mask[B,H,W], ro, pos, nrm, kd, ks [B,H,W,3], view [B,1,1,3], light [Hl,Wl,3], pdf, rows (1-D), cols;
    sel: bool numpy [B,H,W]; bench_fwd: optional (diff, spec) CUDA tensors of the caller's own forward launch with the same seed.
    Returns a dict of counts and relative-L2 errors (see keys below)."""
    import nvdiffrecmc_b200.optixutils as ou
    from nvdiffrecmc_b200.optixutils.ops import env_shade_records
    o = oracle()
    dev = dev_in["ro"].device
    B, H, W = dev_in["ro"].shape[:3]
    S2 = 2 * N * N
    names = ("mask", "ro", "pos", "nrm", "view", "kd", "ks", "light", "pdf", "rows", "cols")
    a = [dev_in[k].detach() for k in names]
    sel_t = torch.tensor(sel, device=dev)

    # ---- product: the records launch (forward + per-ray texel / visibility records) on the FULL launch -----------------
    diff_r, spec_r, rec_t, rec_v = env_shade_records(ctx, *a, perms_dev, BSDF=bsdf, n_samples_x=N, rnd_seed=seed, shadow_scale=shadow_scale,
                                                     batch_offset=batch_offset)
    out = {"pixels_checked": int(sel.sum()), "covered_pixels_checked": int((sel & (dev_in["mask"].cpu().numpy() > 0)).sum())}
    if bench_fwd is not None:
        # the caller's own (non-recording) launch: same rays, same evaluation order
        d0, s0 = bench_fwd
        out["fwd_launch_vs_records_launch_max_abs"] = float(max((d0 - diff_r).abs().max(), (s0 - spec_r).abs().max()))
        diff_p, spec_p = d0, s0
    else:
        diff_p, spec_p = diff_r, spec_r
    g_rec_t = rec_t[sel_t].cpu().numpy(); g_rec_v = rec_v[sel_t].cpu().numpy()
    del rec_t, rec_v
    g_diff = diff_p[sel_t].cpu().numpy(); g_spec = spec_p[sel_t].cpu().numpy()

    # ---- product: forward + backward through the public op, upstream gradients zero outside the selection ---------------
    rng = np.random.default_rng(upstream_seed)
    dg = (rng.uniform(0, 1, size=(B, H, W, 3)) * sel[..., None]).astype(np.float32)
    sg = (rng.uniform(0, 1, size=(B, H, W, 3)) * sel[..., None]).astype(np.float32)
    leaf = {k: dev_in[k].detach().clone().requires_grad_(True) for k in ("pos", "nrm", "kd", "ks", "light")}
    d, s = ou.optix_env_shade(ctx, a[0], a[1], leaf["pos"], leaf["nrm"], a[4], leaf["kd"], leaf["ks"], leaf["light"], a[8], a[9], a[10], BSDF=bsdf,
                              n_samples_x=N, rnd_seed=seed, shadow_scale=shadow_scale, perms=perms_dev, batch_offset=batch_offset)
    torch.autograd.backward([d, s], [torch.tensor(dg, device=dev), torch.tensor(sg, device=dev)])
    g_prod = {k: leaf[k].grad[sel_t].cpu().numpy() for k in ("pos", "nrm", "kd", "ks")}
    outside = ~sel_t
    out["grad_outside_selection_max_abs"] = float(max(leaf[k].grad[outside].abs().max() if bool(outside.any()) else 0.0 for k in ("pos", "nrm", "kd", "ks")))
    lg_prod = leaf["light"].grad.cpu().numpy()
    del leaf, d, s

    # ---- oracle, view by view, mask restricted to the selection ---------------------------------------------------------
    host = {k: dev_in[k].detach().cpu().numpy() for k in names}
    perms = perms_dev.cpu().numpy()
    o_rt = np.full((B, H, W, S2), -1, np.int32); o_rv = np.full((B, H, W, S2), 255, np.uint8)
    o_d = np.zeros((B, H, W, 3), np.float32); o_s = np.zeros((B, H, W, 3), np.float32)
    o_g = {k: np.zeros((B, H, W, 3), np.float32) for k in ("pos", "nrm", "kd", "ks")}
    o_lg = np.zeros_like(host["light"], dtype=np.float64)
    rays = 0
    for b in range(B):
        if not sel[b].any():
            continue
        sl = slice(b, b + 1)
        m = (host["mask"][sl] * sel[sl]).astype(np.float32)
        view = host["view"][sl] if host["view"].shape[0] == B else host["view"]
        args = (scene, m, host["ro"][sl], host["pos"][sl], host["nrm"][sl], view, host["kd"][sl], host["ks"][sl], host["light"], host["pdf"],
                host["rows"], host["cols"], perms)
        kw = dict(BSDF=bsdf, n_samples_x=N, rnd_seed=seed, shadow_scale=shadow_scale, batch_offset=batch_offset + b, vis_mode="bvh")
        dd, ss, (rt, rv) = o.env_shade(*args, records=True, **kw)
        o_d[b], o_s[b], o_rt[b], o_rv[b] = dd[0], ss[0], rt[0], rv[0]
        g = o.env_shade(*args, grads=(dg[sl], sg[sl]), **kw)
        for k, v in zip(("pos", "nrm", "kd", "ks"), g[:4]):
            o_g[k][b] = v[0]
        o_lg += g[4]
        rays += int((m > 0).sum()) * S2
    rt_s, rv_s = o_rt[sel], o_rv[sel]
    cov = host["mask"][sel] > 0
    traced = (g_rec_v != 2) & cov[:, None]
    out["rays_checked"] = int(rays)
    out["texel_mismatch"] = int((g_rec_t[cov] != rt_s[cov]).sum())
    out["vis_mismatch"] = int((g_rec_v[traced] != rv_s[traced]).sum())
    out["rays_traced_checked"] = int(traced.sum())
    out["rays_occluded_checked"] = int((rv_s[traced] == 0).sum())
    rl = {"diff": rel_l2(g_diff, o_d[sel]), "spec": rel_l2(g_spec, o_s[sel]) if bsdf == "pbr" else 0.0}
    for k in ("pos", "nrm", "kd", "ks"):
        ref = o_g[k][sel]
        rl["grad_" + k] = rel_l2(g_prod[k], ref) if np.abs(ref).max() > 0 else float(np.abs(g_prod[k]).max())
    rl["grad_light"] = rel_l2(lg_prod, o_lg) if np.abs(o_lg).max() > 0 else float(np.abs(lg_prod).max())
    out["rel_l2"] = {k: float("%.3e" % v) for k, v in rl.items()}
    out["max_rel_l2"] = float("%.3e" % max(rl.values()))
    return out

"""GPU parity: bilateral denoiser fwd / transposed bwd vs the CPU oracle (denoising.cu restatement)."""
import numpy as np
import pytest
import torch

from common import oracle, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _inputs(B, H, W, seed):
    rng = np.random.default_rng(seed)
    col = rng.uniform(0, 2, size=(B, H, W, 3)).astype(np.float32)
    # piecewise-smooth guides: a few normal "facets" + noise, depth ramp with discontinuities
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    facet = ((xs // 7 + ys // 5) % 3)[None, ..., None]
    n = np.stack([np.sin(facet[..., 0] * 1.3), np.cos(facet[..., 0] * 0.7), np.ones_like(facet[..., 0], dtype=np.float64)], -1)
    n = n + rng.normal(size=(B, H, W, 3)) * 0.05
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    n[:, :2, :3] = 0.0                                    # background pixels have zero normals
    z = (xs * 0.01 + ys * 0.02 + facet[..., 0] * 0.3)[..., None] + rng.normal(size=(B, H, W, 1)) * 0.001
    dz = np.abs(rng.normal(size=(B, H, W, 1))) * 0.02 + 0.001
    return col, n.astype(np.float32), np.concatenate([z, dz], -1).astype(np.float32)


@pytest.mark.parametrize("sigma", [2.0, 0.6, 0.0001])
@pytest.mark.parametrize("shape", [(2, 45, 70), (1, 16, 9)])
def test_forward_backward(dev, sigma, shape):
    import nvdiffrecmc_b200.optixutils as ou
    col, nrm, zdz = _inputs(*shape, seed=int(sigma * 10))
    tc = torch.tensor(col, device=dev, requires_grad=True)
    out = ou.bilateral_denoiser(tc, torch.tensor(nrm, device=dev), torch.tensor(zdz, device=dev), sigma)
    o = oracle()
    raw = o.bilateral_fwd(col, nrm, zdz, sigma)
    assert rel_l2(out.detach().cpu().numpy(), raw[..., :3] / raw[..., 3:]) < TOL
    g = np.random.default_rng(1).uniform(0, 1, size=out.shape).astype(np.float32)
    out.backward(torch.tensor(g, device=dev))
    # autograd of the division (ops.py:141) then the transposed filter; the weight channel gets no gradient (denoising.cu:122)
    og = np.concatenate([g / raw[..., 3:], np.zeros_like(raw[..., 3:])], -1)
    assert rel_l2(tc.grad.cpu().numpy(), o.bilateral_bwd(nrm, zdz, sigma, og)) < TOL


def test_fused_two_signal_path_is_identical(dev):
    import nvdiffrecmc_b200.optixutils as ou
    colA, nrm, zdz = _inputs(2, 40, 50, 3)
    colB = np.random.default_rng(9).uniform(0, 1, size=colA.shape).astype(np.float32)
    a, b, n, z = [torch.tensor(x, device=dev) for x in (colA, colB, nrm, zdz)]
    a.requires_grad_(True); b.requires_grad_(True)
    fa, fb = ou.bilateral_denoiser2(a, b, n, z, 2.0)
    (fa.sum() * 2 + (fb * fb).sum()).backward()
    a2, b2 = a.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    sa, sb = ou.bilateral_denoiser(a2, n, z, 2.0), ou.bilateral_denoiser(b2, n, z, 2.0)
    (sa.sum() * 2 + (sb * sb).sum()).backward()
    assert rel_l2(fa.detach().cpu().numpy(), sa.detach().cpu().numpy()) < 1e-6 and rel_l2(fb.detach().cpu().numpy(), sb.detach().cpu().numpy()) < 1e-6
    assert rel_l2(a.grad.cpu().numpy(), a2.grad.cpu().numpy()) < 1e-6 and rel_l2(b.grad.cpu().numpy(), b2.grad.cpu().numpy()) < 1e-6


def test_module_with_strided_slices(dev):
    """denoiser.forward gets one cat'ed [...,8] tensor and slices it (render.py:120, denoiser.py:27-31)."""
    from nvdiffrecmc_b200.denoiser import BilateralDenoiser
    col, nrm, zdz = _inputs(1, 33, 47, 5)
    x = torch.tensor(np.concatenate([col, nrm * 0.7, zdz], -1), device=dev)
    den = BilateralDenoiser(influence=1.0)
    out = den.forward(x)
    ref = oracle().bilateral_denoiser(col, nrm, zdz, den.sigma)
    assert rel_l2(out.cpu().numpy(), ref) < TOL
    den.set_influence(0.25)
    assert den.sigma == 0.5 and den.N == 2 * 2 + 1


@pytest.mark.parametrize("bsdf", ["pbr", "diffuse"])
def test_denoise_and_combine_matches_the_unfused_tail(dev, bsdf):
    """render.py:119-127: fused (two-signal filter + one combine launch) == the reference's op sequence built from the single ops."""
    import nvdiffrecmc_b200.optixutils as ou
    g = torch.Generator().manual_seed(11)
    B, H, W = 2, 37, 41
    mk = lambda c: torch.rand(B, H, W, c, generator=g).to(dev)
    diff, spec, kd, ks = mk(3), mk(3), mk(3), mk(3)
    nrm = torch.nn.functional.normalize(mk(3) - 0.5, dim=-1)
    zdz = torch.cat([mk(1) + 1.0, torch.full((B, H, W, 1), 0.01, device=dev)], -1)
    ins = [t.clone().requires_grad_(True) for t in (diff, spec, kd, ks)]
    ref_in = [t.clone().requires_grad_(True) for t in (diff, spec, kd, ks)]
    out = ou.denoise_and_combine(ins[0], ins[1], nrm, zdz, 1.0, ins[2], ins[3], BSDF=bsdf)
    d = ou.bilateral_denoiser(ref_in[0], nrm, zdz, 1.0)
    s = ou.bilateral_denoiser(ref_in[1], nrm, zdz, 1.0)
    ref = d * ref_in[2] if bsdf != "pbr" else d * (ref_in[2] * (1.0 - ref_in[3][..., 2:3])) + s
    assert rel_l2(out.detach().cpu().numpy(), ref.detach().cpu().numpy()) < 1e-6
    gout = mk(3)
    out.backward(gout); ref.backward(gout)
    for a, b, name in zip(ins, ref_in, ("diffuse", "specular", "kd", "ks")):
        if bsdf != "pbr" and name in ("specular", "ks"):
            assert a.grad is None or float(a.grad.abs().max()) == 0
            continue
        assert rel_l2(a.grad.cpu().numpy(), b.grad.cpu().numpy()) < 1e-5, name


@pytest.mark.parametrize("sigma", [2.0, 0.6, 0.0001])
@pytest.mark.parametrize("shape", [(2, 45, 72), (1, 130, 256)])
def test_tma_staged_forward(dev, sigma, shape):
    """Contiguous [B,H,W,3] signals / normals, [B,H,W,2] depth and W % 4 == 0 take the TMA-staged forward kernel (denoise.cu:
    bilateral_fwd_tma_kernel: cp.async.bulk.tensor.3d halo tiles, out-of-image taps zero-filled by the copy engine); a channel slice of
    a wider tensor (12-byte pixels at a 16-byte pitch) cannot be described by a tensor map and takes the plain kernel.  Same tap loop:
    the two must agree BIT FOR BIT, and with the oracle to 1e-4, for the single- and the two-signal entry points."""
    import nvdiffrecmc_b200.optixutils as ou
    from nvdiffrecmc_b200.optixutils.ops import _bilateral_denoiser_func, _bilateral_denoiser2_func
    col, nrm, zdz = _inputs(*shape, seed=7 + int(sigma * 10))
    colB = np.random.default_rng(3).uniform(0, 1, size=col.shape).astype(np.float32)
    c, cb, n, z = [torch.tensor(x, device=dev) for x in (col, colB, nrm, zdz)]
    pad = lambda t: torch.cat([t, torch.zeros_like(t[..., :1])], -1)[..., :3]            # same values, pixel pitch 16 B: not TMA-able
    assert not pad(c).is_contiguous()
    raw_tma = _bilateral_denoiser_func.apply(c, n, z, sigma)
    raw_plain = _bilateral_denoiser_func.apply(pad(c), n, z, sigma)
    assert torch.equal(raw_tma, raw_plain)
    a_t, b_t = _bilateral_denoiser2_func.apply(c, cb, n, z, sigma)
    a_p, b_p = _bilateral_denoiser2_func.apply(pad(c), pad(cb), n, z, sigma)
    assert torch.equal(a_t, a_p) and torch.equal(b_t, b_p) and torch.equal(a_t, raw_tma)
    ref = oracle().bilateral_fwd(col, nrm, zdz, sigma)
    assert rel_l2(raw_tma.cpu().numpy(), ref) < TOL
    # transposed filter: contiguous [B,H,W,4] upstream gradients take the TMA kernel (one 128-bit shared load per tap), a strided view the plain one
    from nvdiffrecmc_b200 import _lib as L
    import ctypes as C
    g4 = torch.rand(*shape, 4, device=dev); g4b = torch.rand(*shape, 4, device=dev)
    wide = lambda t: torch.cat([t, torch.zeros_like(t[..., :1])], -1)[..., :4]
    outs = {}
    for tag, (ga, gb) in (("tma", (g4, g4b)), ("plain", (wide(g4), wide(g4b)))):
        ca = torch.empty(*shape, 3, device=dev); cb2 = torch.empty(*shape, 3, device=dev)
        L.check(L.lib().mcs_bilateral_bwd2(C.byref(L.nhwc(n)), C.byref(L.nhwc(z)), float(sigma), C.byref(L.nhwc(ga)), C.byref(L.nhwc(gb)), ca.data_ptr(), cb2.data_ptr(),
                                           L.stream_ptr()), "bilateral_denoiser2 (backward)")
        c1 = torch.empty(*shape, 3, device=dev)
        L.check(L.lib().mcs_bilateral_bwd(C.byref(L.nhwc(n)), C.byref(L.nhwc(z)), float(sigma), C.byref(L.nhwc(ga)), c1.data_ptr(), L.stream_ptr()), "bilateral_denoiser (backward)")
        outs[tag] = (ca, cb2, c1)
    for a, b in zip(outs["tma"], outs["plain"]):
        assert torch.equal(a, b)
    assert torch.equal(outs["tma"][0], outs["tma"][2])
    assert rel_l2(outs["tma"][0].cpu().numpy(), oracle().bilateral_bwd(nrm, zdz, sigma, g4.cpu().numpy())) < TOL

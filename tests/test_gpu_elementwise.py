"""GPU parity: renderutils ops (mcs_*_fwd / _bwd through the reference-shaped Python API) vs the CPU oracle
and vs the package's own PyTorch twin (use_python=True), on the reference tests' input distribution
(torch.rand, renderutils/tests/test_bsdf.py:24-295) incl. every input gradient.  Tolerance: 1e-4 rel-L2
(north_star), in practice ~1e-6."""
import numpy as np
import pytest
import torch

from common import oracle, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _rand(shape, seed, dev):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g).to(dev)


def _check(fn_cuda, fn_orc_fwd, fn_orc_bwd, shapes, dev, out_ch, seed=0, names=None):
    ins = [_rand(s, seed + i, dev).requires_grad_(True) for i, s in enumerate(shapes)]
    out = fn_cuda(*ins)
    dout = _rand(tuple(out.shape), seed + 100, dev)
    out.backward(dout)
    npin = [i.detach().cpu().numpy() for i in ins]
    ref = fn_orc_fwd(*npin)
    assert out.shape[-1] == out_ch
    assert rel_l2(out.detach().cpu().numpy(), ref) < TOL
    gref = fn_orc_bwd(*npin, dout.cpu().numpy())
    for k, (i, g) in enumerate(zip(ins, gref)):
        g = np.asarray(g)
        gi = i.grad.cpu().numpy()
        if gi.shape != g.shape:      # broadcast input: the oracle returns full-grid gradients
            g = g.reshape(-1, g.shape[-1]).sum(0).reshape(gi.shape) if gi.size == g.shape[-1] else g.sum(axis=tuple(d for d in range(3) if gi.shape[d] == 1), keepdims=True)
        e = rel_l2(gi, g)
        assert e < TOL, "grad %s rel-L2 %.3e" % (names[k] if names else k, e)


R = (2, 37, 29)     # ragged: not a multiple of the 4-pixel vector width


def test_pbr_bsdf(dev):
    import nvdiffrecmc_b200.renderutils as ru
    o = oracle()
    for bsdf in ("lambert", "frostbite"):
        _check(lambda *a: ru.pbr_bsdf(*a, bsdf=bsdf), lambda *a: o.pbr_bsdf(*a, bsdf=bsdf), lambda *a: o.pbr_bsdf_bwd(*a, bsdf=bsdf),
               [R + (3,)] * 6, dev, 3, names=["kd", "arm", "pos", "nrm", "view", "light"])


def test_pbr_bsdf_broadcast_view_and_light(dev):
    import nvdiffrecmc_b200.renderutils as ru
    o = oracle()
    shapes = [R + (3,)] * 4 + [(2, 1, 1, 3), (1, 1, 1, 3)]
    _check(lambda *a: ru.pbr_bsdf(*a), lambda *a: o.pbr_bsdf(*a), lambda *a: o.pbr_bsdf_bwd(*a), shapes, dev, 3)


def test_pbr_specular(dev):
    import nvdiffrecmc_b200.renderutils as ru
    o = oracle()
    _check(lambda *a: ru.pbr_specular(*a), lambda *a: o.pbr_specular(*a), lambda *a: o.pbr_specular_bwd(*a), [R + (3,)] * 4 + [R + (1,)], dev, 3)


def test_lambert_frostbite(dev):
    import nvdiffrecmc_b200.renderutils as ru
    o = oracle()
    _check(ru.lambert, o.lambert, o.lambert_bwd, [R + (3,)] * 2, dev, 1)
    _check(ru.frostbite_diffuse, o.frostbite_diffuse, o.frostbite_diffuse_bwd, [R + (3,)] * 3 + [R + (1,)], dev, 1)


def test_primitives(dev):
    import nvdiffrecmc_b200.renderutils as ru
    o = oracle()
    _check(ru._fresnel_shlick, o.fresnel_shlick, o.fresnel_shlick_bwd, [R + (3,), R + (3,), R + (1,)], dev, 3)
    _check(ru._ndf_ggx, o.ndf_ggx, o.ndf_ggx_bwd, [R + (1,)] * 2, dev, 1)
    _check(ru._lambda_ggx, o.lambda_ggx, o.lambda_ggx_bwd, [R + (1,)] * 2, dev, 1)
    _check(ru._masking_smith, o.masking_smith, o.masking_smith_bwd, [R + (1,)] * 3, dev, 1)


@pytest.mark.parametrize("two_sided,opengl", [(True, True), (False, False)])
def test_prepare_shading_normal(dev, two_sided, opengl):
    import nvdiffrecmc_b200.renderutils as ru
    o = oracle()
    _check(lambda *a: ru.prepare_shading_normal(*a, two_sided_shading=two_sided, opengl=opengl),
           lambda *a: o.prepare_shading_normal(*a, two_sided_shading=two_sided, opengl=opengl),
           lambda *a: o.prepare_shading_normal_bwd(*a, two_sided_shading=two_sided, opengl=opengl), [R + (3,)] * 6, dev, 3)
    # perturbed_nrm=None default + broadcast camera position (render.py:99)
    pos, sn, st, gn = [_rand(R + (3,), 10 + i, dev) for i in range(4)]
    view = _rand((2, 1, 1, 3), 20, dev)
    a = ru.prepare_shading_normal(pos, view, None, sn, st, gn)
    b = ru.prepare_shading_normal(pos, view, None, sn, st, gn, use_python=True)
    assert rel_l2(a.cpu().numpy(), b.cpu().numpy()) < TOL


def test_cuda_vs_python_twin(dev):
    """The reference's own test pattern: CUDA op vs use_python=True incl. gradients after an MSE loss."""
    import nvdiffrecmc_b200.renderutils as ru
    for bsdf in ("lambert", "frostbite"):
        a = [_rand((1, 16, 16, 3), 30 + i, dev).requires_grad_(True) for i in range(6)]
        b = [x.detach().clone().requires_grad_(True) for x in a]
        tgt = _rand((1, 16, 16, 3), 40, dev)
        torch.nn.MSELoss()(ru.pbr_bsdf(*a, bsdf=bsdf), tgt).backward()
        torch.nn.MSELoss()(ru.pbr_bsdf(*b, bsdf=bsdf, use_python=True), tgt).backward()
        for x, y in zip(a, b):
            assert rel_l2(x.grad.cpu().numpy(), y.grad.cpu().numpy()) < 1e-3     # fp32 autograd of the twin is the noisier side


def test_strided_inputs(dev):
    import nvdiffrecmc_b200.renderutils as ru
    big = _rand((2, 16, 16, 18), 50, dev)
    parts = [big[..., 3 * i:3 * i + 3] for i in range(6)]
    a = ru.pbr_bsdf(*parts)
    b = ru.pbr_bsdf(*[p.contiguous() for p in parts])
    assert torch.equal(a, b)

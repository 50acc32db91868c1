"""GPU parity, row a17: EnvironmentLight.update_pdf (csrc/light.cu) vs the CPU oracle and the torch-formula twin."""
import numpy as np
import pytest
import torch

from common import oracle, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hw", [(24, 40), (256, 256), (512, 1024), (3, 5), (300, 2049)])
def test_update_pdf_matches_oracle(dev, hw):
    from nvdiffrecmc_b200.light import EnvironmentLight
    g = torch.Generator().manual_seed(hw[0])
    base = torch.rand(hw[0], hw[1], 3, generator=g) ** 4 * 20          # HDR-like dynamic range
    base[hw[0] // 3] = 0                                               # an all-black row: cols stays 0 (light.py:58 guard)
    lgt = EnvironmentLight(base.to(dev))
    pdf, rows, cols = oracle().update_pdf(base.numpy())
    gp, gr, gc = lgt._pdf.cpu().numpy(), lgt.rows[:, 0].cpu().numpy(), lgt.cols.cpu().numpy()
    # fp32 tolerance: the oracle accumulates sequentially in fp32 (error grows with the row length), the kernel in fp64
    assert rel_l2(gp, pdf) < 2e-6 and rel_l2(gc, cols) < 1e-5 and rel_l2(gr, rows) < 1e-5
    o64 = oracle(f64=True).update_pdf(base.numpy().astype(np.float64))
    assert rel_l2(gp, o64[0]) < 2e-7 and rel_l2(gc, o64[2]) < 2e-7 and rel_l2(gr, o64[1]) < 2e-7   # vs the fp64 oracle: rounding only
    assert lgt.rows.shape == hw and lgt._pdf.shape == hw and lgt.cols.shape == hw
    assert abs(float(lgt._pdf.double().sum()) - 1) < 1e-5
    # CDF properties the sampler relies on (kernel.cu:139-160): monotone, ends at exactly 1, black rows all zero
    assert (np.diff(gc, axis=1) >= 0).all() and (np.diff(gr) >= 0).all()
    live = np.ones(hw[0], bool); live[hw[0] // 3] = False
    assert (gc[live, -1] == 1.0).all() and gr[-1] == 1.0 and (gc[~live] == 0).all()
    # torch-formula twin
    tw = EnvironmentLight(base.to(dev)); tw.update_pdf(use_python=True)
    assert rel_l2(gp, tw._pdf.cpu().numpy()) < 2e-6 and rel_l2(gc, tw.cols.cpu().numpy()) < 1e-5 and rel_l2(gr, tw.rows[:, 0].cpu().numpy()) < 1e-5


def test_update_pdf_strided_base_and_determinism(dev):
    from nvdiffrecmc_b200.light import EnvironmentLight
    g = torch.Generator().manual_seed(9)
    big = torch.rand(64, 96, 4, generator=g).to(dev)
    view = big[::2, 8:72, 1:4]                                         # non-contiguous in every dim
    a = EnvironmentLight(view); b = EnvironmentLight(view.contiguous())
    assert torch.equal(a._pdf, b._pdf) and torch.equal(a.cols, b.cols) and torch.equal(a.rows, b.rows)
    a.update_pdf()
    assert torch.equal(a._pdf, b._pdf) and torch.equal(a.cols, b.cols)


def test_env_shade_consumes_native_tables(dev):
    """The tables built on the device drive the sampler: white furnace (constant probe, no occluder in reach) integrates to the
    analytic value -- kernel.cu:403-461 with pdf == light / integral."""
    import nvdiffrecmc_b200.optixutils as ou
    from nvdiffrecmc_b200.light import EnvironmentLight
    from nvdiffrecmc_b200 import synth
    lgt = EnvironmentLight(torch.full((32, 64, 3), 0.7, device=dev))
    ctx = ou.OptiXContext()
    v = torch.tensor([[50, 50, 50], [51, 50, 50], [50, 51, 50]], dtype=torch.float32, device=dev)
    ou.optix_build_bvh(ctx, v, torch.tensor([[0, 1, 2]], dtype=torch.int32, device=dev), rebuild=1)
    B, H, W, N = 1, 8, 8, 8
    pos = torch.zeros(B, H, W, 3, device=dev); nrm = torch.zeros(B, H, W, 3, device=dev); nrm[..., 2] = 1
    view = torch.tensor([0.3, 0.2, 2.0], device=dev).view(1, 1, 1, 3)
    kd = torch.full((B, H, W, 3), 0.5, device=dev); ks = torch.zeros(B, H, W, 3, device=dev); ks[..., 1] = 0.5
    perms = torch.tensor(synth.make_perms(N, seed=1, rows=64), device=dev)
    d, s = ou.optix_env_shade(ctx, torch.ones(B, H, W, device=dev), pos + nrm * 0.001, pos, nrm, view, kd, ks, lgt.base, lgt._pdf, lgt.rows[:, 0],
                              lgt.cols, BSDF='diffuse', n_samples_x=N, rnd_seed=3, perms=perms)
    # diffuse-only: integral of L * cos/pi over the hemisphere = L
    assert abs(float(d.mean()) / 0.7 - 1) < 0.05


def test_update_pdf_matches_reference_golden(dev):
    """The native kernel against the output of the reference's own update_pdf (tests/golden/ref_update_pdf.npz)."""
    import os
    from nvdiffrecmc_b200.light import EnvironmentLight
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_update_pdf.npz"))
    for k in ("a", "b"):
        lgt = EnvironmentLight(torch.tensor(d[k + "_base"], device=dev))
        assert rel_l2(lgt._pdf.cpu().numpy(), d[k + "_pdf"]) < 2e-6 and rel_l2(lgt.cols.cpu().numpy(), d[k + "_cols"]) < 2e-6
        assert rel_l2(lgt.rows.cpu().numpy(), d[k + "_rows"]) < 2e-6 and tuple(lgt.rows.shape) == d[k + "_rows"].shape

"""GPU parity, row f3: fused image loss and batched transforms vs the CPU oracle (and the PyTorch twins)."""
import numpy as np
import pytest
import torch

from common import oracle, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("loss", ["l1", "mse", "smape", "relmse", "n2n"])
@pytest.mark.parametrize("tm", ["none", "log_srgb"])
def test_image_loss(dev, loss, tm):
    import nvdiffrecmc_b200.renderutils as ru
    g = torch.Generator().manual_seed(3)
    img = (torch.rand(2, 37, 29, 3, generator=g) * 3).to(dev); tgt = (torch.rand(2, 37, 29, 3, generator=g) * 3).to(dev)
    img[0, 0, 0] = -0.5; img[0, 0, 1] = 70000.0; tgt[0, 1, 0] = 0.0           # clamped / zero-gradient cases (loss.cu:118-119, 217-222)
    img.requires_grad_(True); tgt.requires_grad_(True)
    v = ru.image_loss(img, tgt, loss=loss, tonemapper=tm)
    v.backward()
    o = oracle()
    ni, nt = img.detach().cpu().numpy(), tgt.detach().cpu().numpy()
    assert abs(float(v) / o.image_loss(ni, nt, loss, tm) - 1) < 1e-5
    gi, gt = o.image_loss_bwd(ni, nt, loss, tm)
    assert rel_l2(img.grad.cpu().numpy(), gi) < 1e-4 and rel_l2(tgt.grad.cpu().numpy(), gt) < 1e-4
    assert float(img.grad[0, 0, 0].abs().max()) == 0 and float(img.grad[0, 0, 1].abs().max()) == 0
    # deterministic reduction
    assert float(ru.image_loss(img, tgt, loss=loss, tonemapper=tm)) == float(v)
    # PyTorch twin on in-range inputs
    a = (torch.rand(1, 16, 16, 3, generator=g) * 2 + 0.01).to(dev); b = (torch.rand(1, 16, 16, 3, generator=g) * 2 + 0.01).to(dev)
    assert abs(float(ru.image_loss(a, b, loss=loss, tonemapper=tm)) / float(ru.image_loss(a, b, loss=loss, tonemapper=tm, use_python=True)) - 1) < 1e-4


@pytest.mark.parametrize("bp", [1, 3])
def test_xfm_points_and_vectors(dev, bp):
    import nvdiffrecmc_b200.renderutils as ru
    g = torch.Generator().manual_seed(4)
    pts = torch.rand(bp, 1001, 3, generator=g).to(dev).requires_grad_(True)
    mtx = torch.rand(3, 4, 4, generator=g).to(dev)
    o = oracle()
    for is_points, fn in ((True, ru.xfm_points), (False, ru.xfm_vectors)):
        pts.grad = None
        out = fn(pts, mtx)
        d = torch.rand(out.shape, generator=g).to(dev)
        out.backward(d)
        ref = o.xfm(pts.detach().cpu().numpy(), mtx.cpu().numpy(), is_points)
        assert out.shape == ref.shape and rel_l2(out.detach().cpu().numpy(), ref) < 1e-6
        gref = o.xfm_bwd(mtx.cpu().numpy(), d.cpu().numpy(), is_points)
        if bp == 1:
            gref = gref.sum(0, keepdims=True)
        assert rel_l2(pts.grad.cpu().numpy(), gref) < 1e-5
        assert rel_l2(out.detach().cpu().numpy(), fn(pts, mtx, use_python=True).detach().cpu().numpy()) < 1e-6

"""Developer tool: bilateral denoiser forward (one / two signals) and backward at 8 x 512^2, sigma = 2; MCS_DENOISE_NO_TMA=1 selects the plain
staging kernel for a same-library A/B.  usage: python tools/dnbench.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import nvdiffrecmc_b200.optixutils as ou
dev = torch.device("cuda:0")
flush = torch.empty(1 << 29, dtype=torch.uint8, device=dev)
def timed(fn, reps=20):
    fn(); fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
g = torch.Generator().manual_seed(0)
B, H, W = 8, 512, 512
col = torch.rand(B, H, W, 3, generator=g).to(dev); colB = torch.rand(B, H, W, 3, generator=g).to(dev)
nrm = torch.nn.functional.normalize(torch.rand(B, H, W, 3, generator=g).to(dev) - 0.5, dim=-1)
zdz = torch.stack([torch.rand(B, H, W, generator=g).to(dev) + 1, torch.full((B, H, W), 0.01, device=dev)], -1)
with torch.no_grad():
    f1 = timed(lambda: ou.bilateral_denoiser(col, nrm, zdz, 2.0))
    f2 = timed(lambda: ou.bilateral_denoiser2(col, colB, nrm, zdz, 2.0))
cg = col.clone().requires_grad_(True); cgb = colB.clone().requires_grad_(True)
from nvdiffrecmc_b200.optixutils.ops import _bilateral_denoiser2_func
ya, yb = _bilateral_denoiser2_func.apply(cg, cgb, nrm, zdz, 2.0)
ga, gb = torch.rand_like(ya), torch.rand_like(yb)
b2 = timed(lambda: torch.autograd.grad([ya, yb], [cg, cgb], [ga, gb], retain_graph=True))
print(json.dumps({"bwd2_ms": round(b2, 4), "path": "plain" if os.environ.get("MCS_DENOISE_NO_TMA") else "tma", "fwd1_ms": round(f1, 4), "fwd2_ms": round(f2, 4),
                  "gtaps_per_s_fwd2": round(B * H * W * 529 / f2 / 1e6, 1)}))

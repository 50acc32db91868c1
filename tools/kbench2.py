"""Developer tool: CUDA-event timings of the secondary kernels (streaming BSDF ops, shading normal, denoiser, BVH build)
against their rooflines (pbr_bsdf 84 B/px fwd, 156 B/px bwd; denoiser 48 B/px; bvh 248 B/tri)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import nvdiffrecmc_b200.renderutils as ru
import nvdiffrecmc_b200.optixutils as ou
from nvdiffrecmc_b200 import synth
import bench

dev = torch.device("cuda:0")
hbm, _ = bench.peaks()


def timeit(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


out = {}
for shape in ((16, 512, 512), (1, 2048, 2048), (1, 256, 256)):          # renderutils/tests/test_perf.py:54-56 + BASELINE config #1
    ins = [torch.rand(*shape, 3, device=dev) for _ in range(6)]
    px = shape[0] * shape[1] * shape[2]
    f = timeit(lambda: ru.pbr_bsdf(*ins))
    insg = [i.clone().requires_grad_(True) for i in ins]
    o = ru.pbr_bsdf(*insg); g = torch.rand_like(o)
    bw = timeit(lambda: torch.autograd.grad(o, insg, g, retain_graph=True))
    pf = timeit(lambda: ru.prepare_shading_normal(*ins))
    out["pbr_bsdf %s" % (shape,)] = {"fwd_ms": round(f, 4), "fwd_GBs": round(84 * px / f / 1e6, 1), "fwd_frac": round(84 * px / f / 1e6 / hbm, 3),
                                     "bwd_ms": round(bw, 4), "bwd_GBs": round(156 * px / bw / 1e6, 1), "bwd_frac": round(156 * px / bw / 1e6 / hbm, 3),
                                     "psn_fwd_ms": round(pf, 4), "psn_fwd_frac": round(84 * px / pf / 1e6 / hbm, 3)}
B, H, W = 8, 512, 512
col = torch.rand(B, H, W, 3, device=dev); col2 = torch.rand(B, H, W, 3, device=dev)
nrm = torch.nn.functional.normalize(torch.rand(B, H, W, 3, device=dev) + 0.5, dim=-1); zdz = torch.rand(B, H, W, 2, device=dev)
t1 = timeit(lambda: ou.bilateral_denoiser(col, nrm, zdz, 2.0))
t2 = timeit(lambda: ou.bilateral_denoiser2(col, col2, nrm, zdz, 2.0))
px = B * H * W
out["bilateral 8x512x512 sigma=2 (529 taps)"] = {"single_ms": round(t1, 3), "fused2_ms": round(t2, 3), "Gtaps_per_s_fused": round(px * 529 * 2 / t2 / 1e6, 1),
                                                  "hbm_frac_single": round(48 * px / t1 / 1e6 / hbm, 4)}
for level, name in ((4, "7168 tris"),):
    v, f = synth.scene_mesh("blob+torus", level=level)
    tv, tf = torch.tensor(v, device=dev), torch.tensor(f, device=dev)
    ctx = ou.OptiXContext()
    t = timeit(lambda: ou.optix_build_bvh(ctx, tv, tf, 1))
    tr = timeit(lambda: ou.optix_build_bvh(ctx, tv, tf, 0))
    out["bvh_build " + name] = {"rebuild_us": round(t * 1e3, 1), "refit_us": round(tr * 1e3, 1)}
# 1M triangles: displaced grid
n = 724
g = np.linspace(-1, 1, n + 1, dtype=np.float32)
X, Z = np.meshgrid(g, g, indexing="ij")
Y = (0.2 * np.sin(7 * X) * np.cos(5 * Z)).astype(np.float32)
v = np.stack([X, Y, Z], -1).reshape(-1, 3)
idx = np.arange((n + 1) * (n + 1)).reshape(n + 1, n + 1)
a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
f = np.concatenate([np.stack([a, c, b], -1).reshape(-1, 3), np.stack([a, d, c], -1).reshape(-1, 3)]).astype(np.int32)
tv, tf = torch.tensor(v, device=dev), torch.tensor(f, device=dev)
ctx = ou.OptiXContext()
t = timeit(lambda: ou.optix_build_bvh(ctx, tv, tf, 1), reps=10, warm=3)
T = f.shape[0]
out["bvh_build %d tris" % T] = {"rebuild_ms": round(t, 3), "GBs_at_248B_per_tri": round(248 * T / t / 1e6, 1), "hbm_frac": round(248 * T / t / 1e6 / hbm, 4)}
ro = torch.rand(2_000_000, 3, device=dev) * 2 - 1; ro[:, 1] = ro[:, 1].abs() + 0.3
rd = torch.nn.functional.normalize(torch.randn(2_000_000, 3, device=dev), dim=-1)
tt = timeit(lambda: ou.trace_visibility(ctx, ro, rd), reps=5, warm=2)
out["trace_visibility 1M tris, 2M incoherent rays"] = {"ms": round(tt, 3), "Mrays_s": round(2.0 / tt * 1e3, 1)}
print(json.dumps(out, indent=1))

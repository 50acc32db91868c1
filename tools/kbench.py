"""Developer tool: time the fused env_shade forward / backward kernels alone on the bench workload.
usage: [MCS_LIB=path/to/variant.so] python tools/kbench.py [views] [res] [n]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
wl = dict(bench.WORKLOAD)
if len(sys.argv) > 1: wl["views_per_gpu"] = int(sys.argv[1])
if len(sys.argv) > 2: wl["res"] = int(sys.argv[2])
if len(sys.argv) > 3: wl["n_samples_x"] = int(sys.argv[3])
dev = torch.device("cuda:0")
w = bench.GpuWorkload(wl, 0, 1, dev)
if "KB_SHADOW" in os.environ:            # e.g. KB_SHADOW=0: skip the trace phase entirely (sampling + shading cost alone)
    _orig = w.ou.optix_env_shade
    w.ou.optix_env_shade = lambda *a, **k: _orig(*a, shadow_scale=float(os.environ["KB_SHADOW"]), **k)
f, b = bench.time_env_kernels(w, reps=5, warm=2)
print(json.dumps({"lib": os.environ.get("MCS_LIB", "default"), "views": wl["views_per_gpu"], "fwd_ms": round(f, 3), "bwd_ms": round(b, 3),
                  "fwd_mrays": round(w.rays_per_pass / f / 1e3, 1), "bwd_mrays": round(w.rays_per_pass / b / 1e3, 1)}))

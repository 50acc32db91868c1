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
if os.environ.get("KB_MESH"): wl["mesh"] = os.environ["KB_MESH"]
if os.environ.get("KB_LEVEL"): wl["mesh_level"] = int(os.environ["KB_LEVEL"])
dev = torch.device("cuda:0")
w = bench.GpuWorkload(wl, 0, 1, dev)
import numpy as np
fw, bw = bench.time_env_kernels(w, reps=int(os.environ.get("KB_REPS", 12)), warm=3, full=True, shadow_scale=float(os.environ.get("KB_SHADOW", 1.0)))   # KB_SHADOW=0: no trace phase
f, b = float(np.median(fw)), float(np.median(bw))
print(json.dumps({"lib": os.path.basename(os.environ.get("MCS_LIB", "default")), "views": wl["views_per_gpu"], "config_key": bench.config_key(wl, wl["views_per_gpu"]),
                  "rays_per_launch": w.rays_per_pass, "fwd_ms": round(f, 3), "fwd_min_max": [round(min(fw), 2), round(max(fw), 2)],
                  "bwd_ms": round(b, 3), "fwd_mrays": round(w.rays_per_pass / f / 1e3, 1), "bwd_mrays": round(w.rays_per_pass / b / 1e3, 1)}))

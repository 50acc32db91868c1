"""Developer tool: one forward + replay-backward of the bench launch with the trainable probe initialised from the HDR environment
(sun texel => light-gradient atomics hot spot) and REAL upstream gradients of the training loss; run under
`ncu --set full -k regex:env_shade_replay`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
wl = dict(bench.WORKLOAD)
dev = torch.device("cuda:0")
w = bench.GpuWorkload(wl, 0, 1, dev, light_init=os.environ.get("LIGHT_INIT", "hdr"))
for _ in range(3):
    w.step()
torch.cuda.synchronize()
print("done", float(w.light_base.max()))

"""Developer tool: capture one bench step in a CUDA graph at a small size and print the full traceback if capture fails;
then check that replays advance the seed (losses differ between replays) and match eager steps from the same state."""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
wl = dict(bench.WORKLOAD, views_per_gpu=int(os.environ.get("GD_VIEWS", 2)), res=int(os.environ.get("GD_RES", 128)))
dev = torch.device("cuda:0")
w = bench.GpuWorkload(wl, 0, 1, dev)
for _ in range(3):
    w.step()
torch.cuda.synchronize()
try:
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        loss = w.step()
    print("capture ok")
except Exception:
    traceback.print_exc()
    sys.exit(1)
# replay vs eager from identical state
state = (w.flat.clone(), w.seed.clone(), [{k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.items()} for st in w.optimizer.state.values()])
losses_g = []
for _ in range(3):
    g.replay(); losses_g.append(float(loss))
flat_g = w.flat.clone()
with torch.no_grad():
    w.flat.copy_(state[0]); w.seed.copy_(state[1])
    for st, saved in zip(w.optimizer.state.values(), state[2]):
        for k, v in saved.items():
            if torch.is_tensor(v):
                st[k].copy_(v)
losses_e = [float(w.step()) for _ in range(3)]
print("graph losses", losses_g); print("eager losses", losses_e)
print("params rel diff", float((w.flat - flat_g).norm() / flat_g.norm()))

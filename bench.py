#!/usr/bin/env python
"""bench.py -- hot-path benchmark (BASELINE.json metric: Mrays/s + train iters/s @512x512, 64 spp).

One "step" = one training iteration's worth of the hot path on one batch of synthetic views:
    update_pdf -> LBVH rebuild -> prepare_shading_normal -> env_shade fwd (2*N^2 shadow rays / covered pixel)
    -> fused bilateral denoise (diffuse + specular) -> recombine + fused log-sRGB L1 image loss -> full backward
    (denoise bwd, env_shade bwd re-tracing all rays, shading-normal bwd, texture scatter)
    -> [N > 1: one NCCL all-reduce over the flat parameter-gradient bucket].
value = shadow rays processed per second over the whole job (fwd + bwd rays, all ranks), in Mrays/s;
a ray is counted the way the reference traces it: covered pixels x 2 N^2 per pass.

Usage (driver contract):  python bench.py --gpus N --steps K --warmup W          (torchrun for N > 1)
                          python bench.py --impl reference ...                    (CPU oracle arm)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOAD = dict(name="spot_metal-like synthetic (BASELINE configs[2]): 8 views x 512x512, n_samples_x=8 (64 spp, 128 shadow rays/px), "
                     "procedural 7168-tri blob+ring mesh, 256x256 trainable probe, bilateral denoiser sigma=2",
                views_per_gpu=8, res=512, n_samples_x=8, mesh="blob+torus", mesh_level=4, light_res=256, tex_res=1024, sigma=2.0)


# ------------------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  nvidia-smi needs a few hundred ms to
    start, so the poller is started before the warm-up steps and every sample is time-stamped; the summary uses the samples that fall
    inside the timed window (and says so), falling back to the samples taken under the warm-up load if the window caught < 2."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.stop, self.th, self.proc = index, [], False, None, None
        self.t_load = self.t0 = self.t1 = None

    def _run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                if self.stop:
                    break
                if line.strip():
                    self.samples.append((time.time(), [x.strip() for x in line.strip().split(",")]))
        except Exception:
            pass

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True); self.th.start(); self.t_load = time.time(); return self

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def close(self):
        self.stop = True
        try:
            self.proc.terminate()
        except Exception:
            pass
        if self.th is not None:
            self.th.join(timeout=6)

    def summary(self):
        inside = [s for t, s in self.samples if self.t0 is not None and self.t0 <= t <= (self.t1 or 1e30)]
        window = "timed region"
        if len(inside) < 2:
            inside = [s for t, s in self.samples if t >= (self.t_load or 0) + 0.5 and t <= (self.t1 or 1e30)]
            window = "warm-up + timed region (timed region shorter than two 50 ms samples)"
        sm = [float(s[0]) for s in inside if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in inside if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in inside if len(s) >= 6 for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm),
                "window": window}


# ------------------------------------------------------------------------------------------------
def build_scene_numpy(wl, rank):
    from nvdiffrecmc_b200 import synth
    if wl["mesh"] == "grid1m":          # BASELINE configs[4]-like: ~1M triangles (displaced height field with an overhang ring)
        n = 724
        g = np.linspace(-1, 1, n + 1, dtype=np.float32)
        X, Z = np.meshgrid(g, g, indexing="ij")
        Y = (0.25 * np.sin(7 * X) * np.cos(5 * Z) - 0.2).astype(np.float32)
        v = np.stack([X, Y, Z], -1).reshape(-1, 3)
        idx = np.arange((n + 1) * (n + 1)).reshape(n + 1, n + 1)
        a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
        f = np.concatenate([np.stack([a, c, b], -1).reshape(-1, 3), np.stack([a, d, c], -1).reshape(-1, 3)]).astype(np.int32)
        v, f = synth.merge((v, f), synth.torus_mesh(R=0.6, r=0.08, nu=256, nv=64, tilt=0.3))
        return v, f, synth.vertex_normals(v, f)
    v, f = synth.scene_mesh(wl["mesh"], level=wl["mesh_level"], seed=5)
    vn = synth.vertex_normals(v, f)
    return v, f, vn


class GpuWorkload:
    """All device-resident state of one rank."""

    def __init__(self, wl, rank, world, dev):
        import torch
        import nvdiffrecmc_b200.optixutils as ou
        from nvdiffrecmc_b200 import synth
        from nvdiffrecmc_b200.light import EnvironmentLight
        from nvdiffrecmc_b200.denoiser import BilateralDenoiser
        self.torch, self.ou, self.wl, self.dev, self.rank, self.world = torch, ou, wl, dev, rank, world
        B, res, N = wl["views_per_gpu"], wl["res"], wl["n_samples_x"]
        v, f, vn = build_scene_numpy(wl, rank)
        self.verts = torch.tensor(v, device=dev); self.tris = torch.tensor(f, device=dev)
        self.ctx = ou.OptiXContext()
        ou.optix_build_bvh(self.ctx, self.verts, self.tris, rebuild=1)
        # G-buffer: primary rays through our own BVH, views = global indices [rank*B, rank*B+B)
        gb = []
        rng = np.random.default_rng(4)
        angs = rng.uniform(0, 2 * np.pi, size=world * B)
        for b in range(B):
            gi = rank * B + b
            mv = synth.orbit_view(angs[gi], tilt=-0.4 + 0.3 * np.sin(gi))
            campos, ro, rd = synth.primary_rays(mv, res)
            tid, tuv = ou.trace_closest(self.ctx, torch.tensor(ro.reshape(-1, 3), device=dev), torch.tensor(rd.reshape(-1, 3), device=dev))
            gb.append(synth.assemble_gbuffer(v, f, vn, tid.cpu().numpy().reshape(res, res), tuv.cpu().numpy().reshape(res, res, 3), campos,
                                             seed=100 + gi))
        st = lambda k: np.stack([g[k] for g in gb])
        self.host = dict(mask=st("mask"), pos=st("pos"), smooth_nrm=st("smooth_nrm"), tangent=st("tangent"), geom_nrm=st("geom_nrm"),
                         view=st("view_pos").reshape(B, 1, 1, 3), depth=st("depth"))
        # texel index of each pixel into the trainable kd / ks textures (stand-in for dr.texture: nearest texel of a hashed uv)
        tr = wl["tex_res"]
        self.host["texel"] = ((rng.integers(0, tr * tr, size=(B, res, res))).astype(np.int64))
        self.pinned = {k: torch.tensor(a).pin_memory() for k, a in self.host.items()}
        self.gb = {k: t.to(dev) for k, t in self.pinned.items()}
        self.covered = int((self.host["mask"] > 0).sum())
        self.rays_per_pass = self.covered * 2 * N * N
        # trainable parameters: light probe + kd / ks textures, gradients live in ONE flat bucket (single all-reduce)
        g = torch.Generator(device="cpu").manual_seed(2)
        n_light, n_tex = wl["light_res"] ** 2 * 3, tr * tr * 3
        self.flat = torch.zeros(n_light + 2 * n_tex, device=dev)
        self.flat_grad = torch.zeros_like(self.flat)
        self.flat[:n_light] = (torch.rand(n_light, generator=g) * 0.5 + 0.25).to(dev)
        self.flat[n_light:n_light + n_tex] = torch.rand(n_tex, generator=g).to(dev)
        ks0 = torch.rand(tr * tr, 3, generator=g); ks0[:, 0] = 0; ks0[:, 1] = 0.1 + 0.9 * ks0[:, 1]
        self.flat[n_light + n_tex:] = ks0.reshape(-1).to(dev)
        self.light_base = self.flat[:n_light].view(wl["light_res"], wl["light_res"], 3).requires_grad_(True)
        self.kd_tex = self.flat[n_light:n_light + n_tex].view(tr * tr, 3).requires_grad_(True)
        self.ks_tex = self.flat[n_light + n_tex:].view(tr * tr, 3).requires_grad_(True)
        self.light_base.grad = self.flat_grad[:n_light].view_as(self.light_base)
        self.kd_tex.grad = self.flat_grad[n_light:n_light + n_tex].view_as(self.kd_tex)
        self.ks_tex.grad = self.flat_grad[n_light + n_tex:].view_as(self.ks_tex)
        # Adam over the flat bucket (train.py:401-409 uses torch.optim.Adam per parameter group; one fused launch here), then the
        # post-step clamps of train.py:455-461 / material ranges of configs/*.json
        self.flat_param = torch.nn.Parameter(self.flat, requires_grad=False)
        self.flat_param.grad = self.flat_grad
        self.optimizer = torch.optim.Adam([self.flat_param], lr=wl.get("lr", 0.01), fused=True)
        self.n_light = n_light
        self.lgt = EnvironmentLight(self.light_base)
        self.denoiser = BilateralDenoiser(influence=wl["sigma"] / 2.0)
        self.perms = torch.tensor(synth.make_perms(N, seed=3), device=dev)
        self.target = torch.rand(B, res, res, 3, generator=g).to(dev)
        self.seed = 0
        self.bytes_h2d = sum(t.numel() * t.element_size() for t in self.pinned.values())
        self.ev = {}

    def upload(self):
        for k, t in self.pinned.items():
            self.gb[k].copy_(t, non_blocking=True)

    def step(self, timers=None):
        """One hot-path training iteration.  Returns the loss tensor (device)."""
        torch, ou, wl = self.torch, self.ou, self.wl
        import nvdiffrecmc_b200.renderutils as ru
        from nvdiffrecmc_b200.raster import texel_fetch
        gb = self.gb
        N = wl["n_samples_x"]
        self.flat_grad.zero_()
        self.lgt.update_pdf()                                                        # train.py:422
        ou.optix_build_bvh(self.ctx, self.verts, self.tris, rebuild=1)               # dlmesh.py:50 (every iteration)
        kd = texel_fetch(self.kd_tex, gb["texel"])                                   # material look-up (stand-in for dr.texture, nearest)
        ks = texel_fetch(self.ks_tex, gb["texel"])
        nrm = ru.prepare_shading_normal(gb["pos"], gb["view"], None, gb["smooth_nrm"], gb["tangent"], gb["geom_nrm"], two_sided_shading=True,
                                        opengl=True)                                 # render.py:99
        ro = gb["pos"] + nrm * 0.001                                                 # render.py:110
        if timers is not None:
            timers["fwd0"].record()
        diff, spec = ou.optix_env_shade(self.ctx, gb["mask"], ro, gb["pos"], nrm, gb["view"], kd, ks, self.lgt.base, self.lgt._pdf,
                                        self.lgt.rows[:, 0], self.lgt.cols, BSDF='pbr', n_samples_x=N, rnd_seed=self.seed, shadow_scale=1.0,
                                        perms=self.perms, batch_offset=self.rank * wl["views_per_gpu"])   # render.py:113-115
        if timers is not None:
            timers["fwd1"].record()
        self.seed += 1
        zdz = torch.stack([gb["depth"], torch.full_like(gb["depth"], 0.01)], -1)
        from nvdiffrecmc_b200.denoiser import _safe_normalize
        shaded = ou.denoise_and_combine(diff, spec, _safe_normalize(nrm), zdz, self.denoiser.sigma, kd, ks)   # render.py:119-127 (+ denoiser.py:28)
        loss = ru.image_loss(shaded, self.target, loss='l1', tonemapper='log_srgb')        # train.py:57-58 ('logl1', the default loss)
        if timers is not None:
            timers["bwd0"].record()
        loss.backward()
        if timers is not None:
            timers["bwd1"].record()
        if self.world > 1:
            torch.distributed.all_reduce(self.flat_grad)                             # ONE collective per step
            self.flat_grad.div_(self.world)
        self.optimizer.step()                                                        # train.py:452
        with torch.no_grad():
            self.flat[:self.n_light].clamp_(min=0.0)                                 # light.clamp_(min=0), train.py:460
            self.flat[self.n_light:].clamp_(0.0, 1.0)                                # material kd / ks ranges
        return loss


def time_env_kernels(w, reps=20, warm=5):
    """CUDA-event time of the fused env_shade forward and backward launches alone (stream = torch current stream): median of `reps`
    after `warm` untimed launches (median of >= 20 after 5 warm-ups); each launch draws a new seed, i.e. new rays."""
    import torch
    import nvdiffrecmc_b200.renderutils as ru
    ou, gb, wl = w.ou, w.gb, w.wl
    N = wl["n_samples_x"]
    with torch.no_grad():
        nrm0 = ru.prepare_shading_normal(gb["pos"], gb["view"], None, gb["smooth_nrm"], gb["tangent"], gb["geom_nrm"])
        ro = gb["pos"] + nrm0 * 0.001
        kd0 = w.kd_tex[gb["texel"]].detach(); ks0 = w.ks_tex[gb["texel"]].detach()
    fw, bw = [], []
    for r in range(reps + warm):
        nrm = nrm0.clone().requires_grad_(True); kd = kd0.clone().requires_grad_(True); ks = ks0.clone().requires_grad_(True)
        light = w.lgt.base.detach().clone().requires_grad_(True)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        d, s = ou.optix_env_shade(w.ctx, gb["mask"], ro, gb["pos"], nrm, gb["view"], kd, ks, light, w.lgt._pdf, w.lgt.rows[:, 0], w.lgt.cols,
                                  n_samples_x=N, rnd_seed=1000 + r, perms=w.perms)
        e[1].record()
        gd, gs = torch.ones_like(d), torch.ones_like(s)
        e[2].record()
        torch.autograd.backward([d, s], [gd, gs])
        e[3].record()
        torch.cuda.synchronize()
        if r >= warm:
            fw.append(e[0].elapsed_time(e[1])); bw.append(e[2].elapsed_time(e[3]))
    return float(np.median(fw)), float(np.median(bw))


def traced_fraction(w, res_s=128):
    """Share of the logical rays (covered px x 2N^2) that the kernel actually traces: rays with n.wi <= 0 contribute exactly zero and
    are skipped (rec_vis == 2).  Measured with the records entry point on view 0 cropped to res_s x res_s."""
    import torch
    import nvdiffrecmc_b200.renderutils as ru
    from nvdiffrecmc_b200.optixutils.ops import env_shade_records
    gb, wl = w.gb, w.wl
    N = wl["n_samples_x"]
    c0 = (wl["res"] - res_s) // 2
    sl = (slice(0, 1), slice(c0, c0 + res_s), slice(c0, c0 + res_s))
    with torch.no_grad():
        nrm = ru.prepare_shading_normal(gb["pos"][sl], gb["view"][0:1], None, gb["smooth_nrm"][sl], gb["tangent"][sl], gb["geom_nrm"][sl])
        ro = gb["pos"][sl] + nrm * 0.001
        kd = w.kd_tex[gb["texel"][sl]].detach(); ks = w.ks_tex[gb["texel"][sl]].detach()
        d, s, rt, rv = env_shade_records(w.ctx, gb["mask"][sl].contiguous(), ro, gb["pos"][sl].contiguous(), nrm, gb["view"][0:1], kd, ks,
                                         w.lgt.base.detach(), w.lgt._pdf, w.lgt.rows[:, 0], w.lgt.cols, w.perms, n_samples_x=N, rnd_seed=0)
    cov = gb["mask"][sl] > 0
    rvc = rv[cov]
    return float((rvc != 2).float().mean()), float((rvc == 1).float().mean())


def traversal_counts(wl, sample_res=64):
    """n_nodes / n_tris per ray of the CANONICAL CPU traversal (oracle LBVH) on a pixel subsample of view 0."""
    from common import oracle
    from nvdiffrecmc_b200 import synth
    o = oracle()
    v, f, vn = build_scene_numpy(wl, 0)
    scene = o.scene(v, f)
    rng = np.random.default_rng(4)
    ang = rng.uniform(0, 2 * np.pi, size=8)[0]
    campos, ro, rd = synth.primary_rays(synth.orbit_view(ang, tilt=-0.4), sample_res)
    tid, tuv = scene.closest_hit(ro.reshape(-1, 3), rd.reshape(-1, 3))
    g = synth.assemble_gbuffer(v, f, vn, tid.reshape(sample_res, sample_res), tuv.reshape(sample_res, sample_res, 3), campos, seed=100)
    view = g["view_pos"].reshape(1, 1, 1, 3)
    nrm = o.prepare_shading_normal(g["pos"][None], view, None, g["smooth_nrm"][None], g["tangent"][None], g["geom_nrm"][None])
    ro_s = g["pos"][None] + nrm * np.float32(0.001)
    light = synth.random_light(wl["light_res"], seed=2)
    pdf, rows, cols = o.update_pdf(light)
    N = wl["n_samples_x"]
    perms = synth.make_perms(N, seed=3, rows=1024)
    t0 = time.time()
    d, s, cnt = o.env_shade(scene, g["mask"][None], ro_s, g["pos"][None], nrm, view, g["kd"][None], g["ks"][None], light, pdf, rows, cols, perms,
                            n_samples_x=N, rnd_seed=0, vis_mode="bvh", counters=True)
    dt = time.time() - t0
    rays = float(cnt[0])
    return dict(nodes_per_ray=float(cnt[1]) / rays, tris_per_ray=float(cnt[2]) / rays, rays=int(rays), seconds=dt,
                sample="view 0 at %dx%d, n_samples_x=%d" % (sample_res, sample_res, N))


# ------------------------------------------------------------------------------------------------
def cpu_reference(o):
    """(object with env_shade / bilateral_fwd / bilateral_bwd, kind): the reference's own raygen program and denoiser kernels compiled for
    the host (oracle/_ref, built where /root/reference exists and shipped as a prebuilt library) when available, else the oracle port."""
    try:
        from oracle import Reference
        return Reference(o), "reference"
    except Exception:
        return o, "port"


def cpu_reference_step(o, case, N, sigma, seed, env_shade=None):
    """The same hot-path step on the host CPUs (all OpenMP threads): env_shade forward + backward through `env_shade` (the compiled
    reference or the oracle port), LBVH rebuild / shading normal / denoiser / update_pdf through the oracle port."""
    c = case
    impl = env_shade if env_shade is not None else o
    es = impl.env_shade
    scene = o.scene(c["verts"], c["tris"])                                       # LBVH rebuild every iteration
    nrm = o.prepare_shading_normal(c["pos"], c["view"], None, c["smooth_nrm"], c["tangent"], c["geom_nrm"])
    ro = (c["pos"] + nrm * np.float32(0.001)).astype(np.float32)
    pdf, rows, cols = o.update_pdf(c["light"])
    args = (scene, c["mask"], ro, c["pos"], nrm, c["view"], c["kd"], c["ks"], c["light"], pdf, rows, cols, c["perms"])
    d, s = es(*args, n_samples_x=N, rnd_seed=seed, vis_mode="bvh")
    zdz = np.stack([c["depth"], np.full_like(c["depth"], 0.01)], -1)
    nn = nrm / np.maximum(np.linalg.norm(nrm, axis=-1, keepdims=True), 1e-20)
    fd, fs = impl.bilateral_fwd(d, nn, zdz, sigma), impl.bilateral_fwd(s, nn, zdz, sigma)
    gd = np.concatenate([np.ones_like(d) / fd[..., 3:], np.zeros_like(fd[..., 3:])], -1)
    cd, cs = impl.bilateral_bwd(nn, zdz, sigma, gd), impl.bilateral_bwd(nn, zdz, sigma, gd)
    if impl is o:
        g = es(*args, n_samples_x=N, rnd_seed=seed, vis_mode="bvh", grads=(cd, cs), parallel_bwd=True)
    else:
        g = es(*args, n_samples_x=N, rnd_seed=seed, vis_mode="bvh", grads=(cd, cs))
    o.prepare_shading_normal_bwd(c["pos"], c["view"], None, c["smooth_nrm"], c["tangent"], c["geom_nrm"], g[1])
    return float(d.sum())


def host_cores():
    """CPU threads this process may actually use (cgroup / affinity aware), also exported to OpenMP."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:                                   # cgroup v2 cpu quota
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(per) + 0.5)))
    except Exception:
        pass
    os.environ["OMP_NUM_THREADS"] = str(n)          # torchrun exports 1; the oracle legs also call Oracle.set_threads(n)
    return n


REF_ARM_NOTE = {
    "reference": "env_shade forward + backward and the bilateral denoiser forward + backward = the reference's own kernels (render/optixutils/c_src/"
                 "envsampling/kernel.cu with bsdf.h, math_utils.h; denoising.cu) compiled for the host cores (oracle/_ref, OpenMP over pixels), shadow "
                 "rays answered by the oracle's LBVH (OptiX itself is closed source and needs an RT driver); shading normal / update_pdf / LBVH build "
                 "= oracle C port",
    "port": "CPU oracle port of kernel.cu/denoising.cu/normal.cu (oracle/_ref not available on this machine; OptiX needs libnvoptix + RT driver)",
}


def physical_limiter(path=None):
    """What the committed ncu capture of the dominant kernel says actually bounds it (the contract's HBM roofline counts logical bytes that
    this L1/L2-resident workload never moves): issue-slot utilisation, lanes per instruction, L1 data-pipe and DRAM utilisation."""
    path = path or os.path.join(ROOT, "profiles", "r01_v6_envshade_summary.json")
    try:
        with open(path) as f:
            k = json.load(f)["kernels"][0]
        num = lambda key: float(str(k[key]).split()[0])
        return {"limiter": "instruction issue", "source": os.path.relpath(path, ROOT),
                "issue_active_pct_of_peak": round(num("smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
                "active_lanes_per_instruction": round(num("smsp__thread_inst_executed_per_inst_executed.ratio"), 1),
                "l1_data_pipe_pct_of_peak": round(num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"), 1),
                "dram_pct_of_peak": round(num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), 2)}
    except Exception:
        return None


def run_reference(args, wl):
    """--impl reference: the reference's algorithm on the host CPUs (oracle port; OptiX cannot be built/run here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    from common import make_case, oracle
    o = oracle()
    cores = o.set_threads(cores)
    es, kind = cpu_reference(o)
    N, res_s = wl["n_samples_x"], 128
    case = make_case(res=res_s, B=1, N=N, mesh=wl["mesh"], level=wl["mesh_level"], light="random", light_hw=(wl["light_res"], wl["light_res"]),
                     perm_rows=4096)
    covered = int((case["mask"] > 0).sum())
    rays_step = covered * 2 * N * N * 2
    for i in range(args.warmup):
        cpu_reference_step(o, case, N, wl["sigma"], i, es)
    t0 = time.time()
    for i in range(args.steps):
        cpu_reference_step(o, case, N, wl["sigma"], 100 + i, es)
    dt = (time.time() - t0) / args.steps
    val = rays_step / dt / 1e6
    sample = "1 view at %dx%d of the same scene (same mesh, probe, n_samples_x=%d, sigma=%g): %d rays/step" % (res_s, res_s, N, wl["sigma"], rays_step)
    emit({
        "impl": "reference", "metric": "shadow_rays_per_second_train_step", "value": round(val, 4), "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "reference_arm": REF_ARM_NOTE[kind]},
        "cpu_baseline": {"value": round(val, 4), "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(val, 4), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "iters_per_s_on_sample": round(1.0 / dt, 4),
    })


# ------------------------------------------------------------------------------------------------
_STDOUT_FD = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner when NCCL_DEBUG is set, torchrun
    notices), so everything but the final line is routed to stderr at the file-descriptor level."""
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(obj):
    sys.stdout.flush()
    if _STDOUT_FD is not None:
        os.dup2(_STDOUT_FD, 1)
    print(json.dumps(obj))
    sys.stdout.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mcshade", choices=["mcshade", "reference"])
    ap.add_argument("--views", type=int, default=None, help="override views per GPU (debug)")
    ap.add_argument("--res", type=int, default=None, help="override resolution (debug)")
    ap.add_argument("--n", type=int, default=None, help="override n_samples_x (debug)")
    ap.add_argument("--mesh", default=None, help="override mesh: blob | blob+torus | full | grid1m (debug)")
    ap.add_argument("--level", type=int, default=None, help="override icosphere subdivision level (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    quiet_stdout()
    wl = dict(WORKLOAD)
    if args.views:
        wl["views_per_gpu"] = args.views
    if args.res:
        wl["res"] = args.res
    if args.n:
        wl["n_samples_x"] = args.n
    if args.mesh:
        wl["mesh"] = args.mesh
    if args.level is not None:
        wl["mesh_level"] = args.level
    if args.views or args.res or args.n or args.mesh or args.level is not None:
        wl["name"] = "DEBUG OVERRIDE of " + wl["name"]
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    import torch.distributed as dist
    from nvdiffrecmc_b200 import _lib
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback for the product path)"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()

    w = GpuWorkload(wl, rank, world, dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (value) --------------------------------------------------
    clk = ClockSampler(local).start()
    for _ in range(args.warmup):
        w.step()
    barrier()
    names = ["fwd0", "fwd1", "bwd0", "bwd1"]
    timers = [{n: torch.cuda.Event(enable_timing=True) for n in names} for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.LAUNCHES.clear()
    barrier()
    clk.begin()
    e0.record()
    for i in range(args.steps):
        w.step(timers[i])
    e1.record()
    barrier()
    clk.end()
    clk.close()
    launches = sum(_lib.LAUNCHES.values())
    ms = e0.elapsed_time(e1)
    t_ms = torch.tensor([ms], device=dev)
    rays = torch.tensor([float(w.rays_per_pass)], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(rays, op=dist.ReduceOp.SUM)
    ms_step = float(t_ms) / args.steps
    total_rays_step = float(rays) * 2.0               # forward + backward re-trace
    value = total_rays_step / (ms_step * 1e-3) / 1e6
    fwd_ms = float(np.median([t["fwd0"].elapsed_time(t["fwd1"]) for t in timers]))
    bwd_all_ms = float(np.median([t["bwd0"].elapsed_time(t["bwd1"]) for t in timers]))

    # ---- end to end through the public API with HOST inputs (e2e) ----------------------------
    # Every step uploads ITS G-buffer from pinned host memory and reads ITS loss + parameter gradients back, all inside the timed
    # region.  The upload of step i+1 runs on a copy stream while step i computes (two device buffer sets, event-ordered), the
    # way a training loop with a prefetching data loader behaves.
    copy_stream = torch.cuda.Stream(device=dev)
    sets = [w.gb, {k: torch.empty_like(t) for k, t in w.gb.items()}]
    up_done = [torch.cuda.Event() for _ in range(2)]
    use_done = [torch.cuda.Event() for _ in range(2)]
    host_out = torch.empty(w.flat_grad.numel() + 1, pin_memory=True)

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(use_done[slot])          # the step that last used this buffer set has finished
            for k, t in w.pinned.items():
                sets[slot][k].copy_(t, non_blocking=True)
            up_done[slot].record(copy_stream)

    def e2e_loop(n):
        for i in range(2):
            use_done[i].record()
        upload(0)
        for i in range(n):
            slot = i & 1
            if i + 1 < n:
                upload(slot ^ 1)                             # prefetch the next step's inputs
            torch.cuda.current_stream().wait_event(up_done[slot])
            w.gb = sets[slot]
            loss = w.step()
            use_done[slot].record()
            host_out[:1].copy_(loss.detach().reshape(1), non_blocking=True)      # device -> host: loss + parameter gradients
            host_out[1:].copy_(w.flat_grad, non_blocking=True)

    e2e_loop(2)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    e2e_loop(args.steps)
    e3.record()
    barrier()
    w.gb = sets[0]
    t2 = torch.tensor([e2.elapsed_time(e3)], device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms_step = float(t2) / args.steps
    e2e_value = total_rays_step / (e2e_ms_step * 1e-3) / 1e6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (rank 0) --------------------------------------------
    k_fwd_ms, k_bwd_ms = time_env_kernels(w)
    hbm, hbm_src = peaks()
    N = wl["n_samples_x"]
    tc = traversal_counts(wl)
    tfrac, vfrac = traced_fraction(w)
    vfrac_pre = vfrac
    # SURVEY 8d:  A_ray = P/(2N^2) + 4 [perms] + 16 [light texel + pdf] + 44 [CDF probes] + 32*nodes + 36*tris   (+12 B/ray light-grad atomics in bwd)
    a_fwd = 88.0 / (2 * N * N) + 4 + 16 + 44 + 32 * tc["nodes_per_ray"] + 36 * tc["tris_per_ray"]
    # backward replays the forward RAY RECORD: per covered pixel 136 B (G-buffer + upstream + gradients), per EVALUATED ray
    # 20 B record + 12 B env texel + 12 B gradient atomics; nothing for sampling or traversal.  Expressed per logical ray:
    a_bwd = 136.0 / (2 * N * N) + vfrac_pre * (20 + 12 + 12)
    ach_fwd = a_fwd * w.rays_per_pass / (k_fwd_ms * 1e-3) / 1e9
    ach_bwd = a_bwd * w.rays_per_pass / (k_bwd_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "env_shade_kernel<0> (fused env sampling + shadow rays + BSDF, forward)",
            "achieved": round(ach_fwd, 2), "peak": hbm, "unit": "GB/s", "frac": round(ach_fwd / hbm, 4), "traffic": None,
            "peak_source": hbm_src, "algorithmic_bytes_per_ray": round(a_fwd, 1), "rays_per_launch": w.rays_per_pass,
            "kernel_ms": round(k_fwd_ms, 3), "mrays_per_s": round(w.rays_per_pass / k_fwd_ms / 1e3, 1),
            "canonical_traversal": tc,
            "traced_fraction": round(tfrac, 4), "visible_fraction": round(vfrac, 4),
            "frac_traced_rays_only": round(ach_fwd * tfrac / hbm, 4),
            "backward": {"kernel": "env_shade_replay_kernel (ray-record replay: adjoint BSDF + gradient scatter only; no sampling, no traversal)", "achieved": round(ach_bwd, 2), "frac": round(ach_bwd / hbm, 4), "kernel_ms": round(k_bwd_ms, 3),
                         "algorithmic_bytes_per_ray": round(a_bwd, 1), "mrays_per_s": round(w.rays_per_pass / k_bwd_ms / 1e3, 1)},
            "note": "achieved = LOGICAL bytes (SURVEY 8d model: every CDF probe, texel, canonical-LBVH node and triangle counted as a memory access) "
                    "x logical rays / kernel time; all tables of this mesh are L1/L2 resident so the physical DRAM traffic (`traffic`) is ~%.1f B/ray "
                    "and the kernel is instruction-issue bound (`physical_limiter`, profiles/r01_v6_*). Rays with n.wi<=0 (exactly zero contribution) are "
                    "counted as the reference counts them but not traced: see traced_fraction / frac_traced_rays_only. The backward kernel "
                    "replays the forward hit record instead of tracing." % (88.0 / (2 * N * N) + 4)}
    prof = os.path.join(ROOT, "profiles", "r01_env_shade_fwd_traffic.json")
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                roof["traffic"] = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
    roof["physical_limiter"] = physical_limiter()

    cpu = None
    if not args.no_cpu_baseline:
        ncores = host_cores()
        from common import make_case, oracle
        o = oracle()
        ncores = o.set_threads(ncores)
        es, kind = cpu_reference(o)
        res_s = 96
        case = make_case(res=res_s, B=1, N=N, mesh=wl["mesh"], level=wl["mesh_level"], light="random", light_hw=(wl["light_res"], wl["light_res"]),
                         perm_rows=4096)
        cov = int((case["mask"] > 0).sum())
        cpu_reference_step(o, case, N, wl["sigma"], 0, es)
        t0 = time.time(); reps = 2
        for i in range(reps):
            cpu_reference_step(o, case, N, wl["sigma"], 1 + i, es)
        dt = (time.time() - t0) / reps
        cpu = {"value": round(cov * 2 * N * N * 2 / dt / 1e6, 4), "unit": "Mrays/s", "cores": ncores, "kind": kind,
               "sample": "same step on 1 view at %dx%d (%d rays/step), OpenMP; %s" % (res_s, res_s, cov * 2 * N * N * 2, REF_ARM_NOTE[kind])}

    out = {
        "metric": "shadow_rays_per_second_train_step", "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "views_per_gpu": wl["views_per_gpu"], "global_views": wl["views_per_gpu"] * world, "res": wl["res"],
                   "n_samples_x": N, "rays_per_covered_pixel": 2 * N * N, "covered_pixels_rank0": w.covered,
                   "coverage_rank0": round(w.covered / (wl["views_per_gpu"] * wl["res"] ** 2), 4),
                   "parallelism": "dp%d over views, one NCCL all-reduce of the flat gradient bucket (%.1f MB)" % (world, w.flat_grad.numel() * 4 / 1e6),
                   "l2_policy": "per-step inputs (G-buffer %.0f MB + intermediates) exceed the 126 MB L2" % (w.bytes_h2d / 1e6)},
        "rays_counted": "covered px x 2N^2 per pass x 2 passes: the reference traces forward AND backward; here forward traces and records the evaluated rays (20 B/sample slot), backward replays the record",
        "ray_record_bytes": int(wl["views_per_gpu"] * wl["res"] ** 2 * (2 * N * N * 20 + 4)),
        "train_iters_per_s": round(1e3 / ms_step, 3),
        "breakdown_ms": {"env_shade_fwd": round(fwd_ms, 3), "backward_all": round(bwd_all_ms, 3), "env_shade_fwd_kernel": round(k_fwd_ms, 3),
                         "env_shade_bwd_kernel": round(k_bwd_ms, 3)},
        "e2e": {"value": round(e2e_value, 2), "unit": "Mrays/s", "ms_per_step": round(e2e_ms_step, 3), "h2d_bytes_per_step": int(w.bytes_h2d),
                "d2h_bytes_per_step": int(host_out.numel() * 4), "train_iters_per_s": round(1e3 / e2e_ms_step, 3)},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
        "roofline": roof,
        "cpu_baseline": cpu,
    }
    if rank == 0:
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""CPU: the bench.py output contract on the leg that runs without a GPU (`--impl reference` = the oracle port on the host cores):
exactly ONE line on stdout, valid JSON, with the keys the driver reads; everything else (library banners, progress) goes to stderr."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")                 # what torchrun exports; the arm must still use all host cores
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    from oracle import build_ref
    assert d["cpu_baseline"]["kind"] == ("reference" if build_ref() else "port")      # oracle/_ref is used whenever it exists
    assert d["cpu_baseline"]["cores"] >= 1 and d["e2e"]["h2d_bytes_per_step"] == 0
    import bench
    assert d["cpu_baseline"]["cores"] == bench.host_cores()     # not the 1 thread of the inherited OMP_NUM_THREADS
    # the reference arm reports on the PRODUCT arm's config (both lines build it with bench.make_config); the bounded sample it actually
    # times is stated inside that config and, with its ray count, in cpu_baseline.sample / reference_sample
    assert d["config"] == bench.make_config(dict(bench.WORKLOAD), bench.WORKLOAD["views_per_gpu"], 1)
    assert "BOUNDED SAMPLE" in d["config"]["reference_arm_sampling"] and "BOUNDED SAMPLE" in d["reference_sample"]


def test_non_zero_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=120, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_dominant_kernel_profile_is_per_config():
    """The physical roofline numbers (thread-instructions, DRAM bytes per launch) come from a committed ncu capture OF THE SAME CONFIG
    (profiles/r02_dominant_kernel.json); any other configuration gets None -- never another config's constant (`traffic` was
    the 512^2 number on the 800^2 and 1 M-triangle lines)."""
    import bench
    wl = dict(bench.WORKLOAD)
    assert bench.config_key(wl, 8) == "8x512x512_n8_blob+torus4_light256"
    assert bench.dominant_kernel_profile(dict(wl, res=640), 8) is None          # no capture committed for this configuration
    assert bench.dominant_kernel_profile(dict(wl, n_samples_x=4), 8) is None
    seen = {}
    for w_, views in ((wl, 8), (dict(wl, res=800), 8), (dict(wl, res=1024, n_samples_x=16, mesh="grid1m", mesh_level=0), 1)):
        prof = bench.dominant_kernel_profile(w_, views)
        if prof is not None:
            for k in ("thread_inst_per_launch", "dram_bytes_per_launch", "rays_per_launch", "source"):
                assert k in prof, k
            assert prof["source"].startswith("profiles/")
            seen[bench.config_key(w_, views)] = prof["dram_bytes_per_launch"]
    assert len(set(seen.values())) == len(seen), "every configuration carries ITS OWN capture"


def test_coverage_balanced_deal():
    import bench
    import numpy as np
    cover = np.random.default_rng(0).integers(60000, 110000, size=64)
    order = bench.deal_views(cover, 8, 8)
    assert sorted(order) == list(range(64))
    per_rank = [int(sum(cover[g] for g in order[r * 8:(r + 1) * 8])) for r in range(8)]
    assert (max(per_rank) - min(per_rank)) / np.mean(per_rank) < 0.01          # rays per rank within 1 %
    assert bench.deal_views(cover[:8], 1, 8) == list(np.argsort(-cover[:8], kind="stable"))

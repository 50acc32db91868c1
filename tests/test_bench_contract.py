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


def test_non_zero_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=120, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_physical_limiter_is_read_from_the_committed_capture():
    import bench
    p = bench.physical_limiter()
    assert p is not None and p["limiter"] == "instruction issue" and p["source"].startswith("profiles/")
    assert 50 < p["issue_active_pct_of_peak"] <= 100 and 1 <= p["active_lanes_per_instruction"] <= 32 and p["dram_pct_of_peak"] < 50
    assert bench.physical_limiter("/nonexistent.json") is None          # a missing capture never breaks the bench line

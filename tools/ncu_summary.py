"""Developer tool: condense an .ncu-rep (ncu --set full) into the per-kernel JSON summaries kept under profiles/.
usage: python tools/ncu_summary.py <report.ncu-rep> <out.json> "<capture description>"
       python tools/ncu_summary.py --launches <launches.csv> <out.json> "<command>"      (gpu__time_duration launch list -> shares)"""
import csv, json, subprocess, sys, collections

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def full(rep, out, desc):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        k = {"Kernel Name": d["Kernel Name"]}
        for key in KEYS:
            if key in d and d[key] != "":
                k[key] = ("%s %s" % (d[key], u[key])).strip()
        kernels.append(k)
    json.dump({"capture": desc, "kernels": kernels}, open(out, "w"), indent=1)
    print("wrote", out, "(%d kernels)" % len(kernels))


def launches(path, out, desc):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.defaultdict(float); cnt = collections.Counter()
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        name = r[ik].split("(")[0][:80]
        tot[name] += v; cnt[name] += 1
    total = sum(tot.values())
    top = [{"kernel": k, "total_ms": round(v, 3), "share_pct": round(100 * v / total, 2), "launches": cnt[k], "avg_ms": round(v / cnt[k], 4)}
           for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:16]]
    json.dump({"command": desc, "total_kernel_ms": round(total, 2), "top": top}, open(out, "w"), indent=1)
    print("wrote", out)


def dominant(rep, kbench_log, out, source):
    """profiles/r02_dominant_kernel.json: per-config physical numbers of env_shade_kernel<0> that bench.py's roofline reads."""
    import os
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    d = dict(zip(rows[0], rows[2])); u = dict(zip(rows[0], rows[1]))
    kb = json.loads([l for l in open(kbench_log) if l.startswith("{")][-1])
    num = lambda k: float(d[k].replace(",", ""))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    dram = num("dram__bytes_read.sum") * scale[u["dram__bytes_read.sum"]] + num("dram__bytes_write.sum") * scale[u["dram__bytes_write.sum"]]
    winst, lanes = num("smsp__inst_executed.sum"), num("smsp__thread_inst_executed_per_inst_executed.ratio")
    entry = {"kernel": d["Kernel Name"], "rays_per_launch": kb["rays_per_launch"], "warp_inst_per_launch": int(winst), "lanes_per_inst": lanes,
             "thread_inst_per_launch": int(winst * lanes), "dram_bytes_per_launch": int(dram), "kernel_ms_under_ncu": num("gpu__time_duration.sum"),
             "issue_active_pct": round(num("smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
             "alu_pipe_pct": round(num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), 1),
             "fma_pipe_pct": round(num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), 1),
             "xu_pipe_pct": round(num("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"), 1),
             "l1_hit_pct": round(num("l1tex__t_sector_hit_rate.pct"), 1), "source": source}
    allc = json.load(open(out)) if os.path.exists(out) else {"what": "ncu --set full --clock-control none capture of env_shade_kernel<0> per bench configuration "
                                                                      "(tools/ncu_env.sh); bench.py's roofline reads the entry of ITS configuration or reports null", "configs": {}}
    allc["configs"][kb["config_key"]] = entry
    json.dump(allc, open(out, "w"), indent=1)
    print("wrote", out, kb["config_key"])


if __name__ == "__main__":
    if sys.argv[1] == "--dominant":
        dominant(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5])
    elif sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[1], sys.argv[2], sys.argv[3])

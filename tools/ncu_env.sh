#!/bin/bash
# usage (on the GPU box): [KB_MESH=.. KB_LEVEL=..] tools/ncu_env.sh <variant-name|default> <out-prefix> [views] [res] [n_samples_x] -- one `ncu --set full` capture of env_shade_kernel<0> on the bench workload
set -e
v=$1; out=$2; views=${3:-8}; res=${4:-512}; n=${5:-8}
if [ "$v" != "default" ]; then export MCS_LIB=nvdiffrecmc_b200/lib/variants/$v.so; fi
KB_REPS=1 ncu --set full --clock-control none --import-source on -k regex:env_shade_kernel -s 3 -c 1 -f -o gpurun_out/$out python tools/kbench.py $views $res $n > gpurun_out/$out.log 2>&1
ncu -i gpurun_out/$out.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
keys=['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','dram__bytes_read.sum','dram__bytes_write.sum','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers']
for r in rows[2:]:
    d=dict(zip(h,r)); print(d['Kernel Name'][:50])
    for k in keys: print('  ',k, d.get(k))
"

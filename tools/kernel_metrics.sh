#!/bin/bash
# usage (under gpurun): tools/kernel_metrics.sh <kind> <messages> <kernel-regex> <tree-dir>...   a few ncu counters of the kernels that match,
# third pass of tools/profile_run.py, per source tree: instructions, issue utilisation, lane use, the main stall reasons
kind=$1; n=$2; k=$3; shift 3
root=$(pwd)
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_drain_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct
for t in "$@"; do
  tag=$(basename $(cd $t && pwd))
  ( cd $t && ncu --metrics $M --clock-control none -k regex:$k --csv --log-file $root/gpurun_out/km_$tag.csv python tools/profile_run.py $kind $n 3 > $root/gpurun_out/km_$tag.log 2>&1 )
  python - "$root/gpurun_out/km_$tag.csv" "$tag" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
ids = sorted({int(r[0]) for r in rows})
names = {int(r[0]): r[4] for r in rows}
# last launch of each distinct kernel name
last = {}
for i in ids: last[names[i]] = i
for nm, i in last.items():
    print(sys.argv[2], nm.split('(')[0])
    for r in rows:
        if int(r[0]) == i: print('   %-88s %s' % (r[-3], r[-1]))
PY
done

#!/bin/bash
# usage (under gpurun): tools/kernel_times.sh <kind> <messages> <tree-dir>...   per-kernel durations (ncu launch list, last of 3 passes) of
# one device-resident pass over <messages> corpus messages, for each source tree
kind=$1; n=$2; shift 2
root=$(pwd)
for t in "$@"; do
  tag=$(basename $(cd $t && pwd))
  ( cd $t && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $root/gpurun_out/kt_$tag.csv python tools/profile_run.py $kind $n 3 > $root/gpurun_out/kt_$tag.log 2>&1 )
  python - "$root/gpurun_out/kt_$tag.csv" "$tag" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
names = [r[4] for r in rows]; vals = [float(r[-1].replace(',', '')) for r in rows]; units = [r[-2] for r in rows]
per = len(rows) // 3
out = []
for nm, v, u in list(zip(names, vals, units))[-per:]:
    us = v / 1000 if u.startswith('ns') or u == 'nsecond' else (v if u.startswith('us') else v * 1000)
    out.append((nm.split('(')[0][-40:], round(us, 1)))
print(sys.argv[2], 'total_us', round(sum(x[1] for x in out), 1), out)
PY
done

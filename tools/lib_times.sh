#!/bin/bash
# usage (under gpurun): tools/lib_times.sh <kind> <messages> <variant.so|default>...   per-kernel durations of one device-resident pass
# for variant libraries of the working tree (SDB200_LIB), like tools/kernel_times.sh does for source trees
kind=$1; n=$2; shift 2
for lib in "$@"; do
  if [ "$lib" = default ]; then unset SDB200_LIB; else export SDB200_LIB=$lib; fi
  tag=$(basename "$lib" .so)
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lt_$tag.csv python tools/profile_run.py $kind $n 3 > gpurun_out/lt_$tag.log 2>&1
  python - gpurun_out/lt_$tag.csv "$tag" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
per = len(rows) // 3
out = [(r[4].split('(')[0][-32:], round(float(r[-1].replace(',', '')) / 1000, 1)) for r in rows[-per:]]
print(sys.argv[2], 'total_us', round(sum(x[1] for x in out), 1), [x for x in out if x[1] > 20])
PY
done
unset SDB200_LIB

#!/bin/bash
# usage (under gpurun): tools/tree_bench.sh <messages> <tree-dir>...   one bench line per source tree under build_variants/
# (each a full copy of the repo at some state, library built in place): mixed / e2e / per-class M msg/s
n=$1; shift
root=$(pwd)
for t in "$@"; do
  tag=$(basename $t)
  ( cd $t && python bench.py --messages $n --steps 3 --warmup 3 --no-cpu --no-lines > $root/gpurun_out/tree_$tag.json 2> $root/gpurun_out/tree_$tag.err || tail -3 $root/gpurun_out/tree_$tag.err
    python -c "
import json;d=json.load(open('$root/gpurun_out/tree_$tag.json'));print('$tag','mixed',round(d['value']/1e6,2),'e2e',round(d['e2e']['value']/1e6,2),{k:round(v['msgs_per_s']/1e6,2) for k,v in d['per_kernel'].items()})" )
done

#!/bin/bash
# usage (under gpurun): tools/variant_bench.sh <messages> <variant.so>...   -> one line per library: mixed / e2e / per-class M msg/s
n=$1; shift
for lib in default "$@"; do
  if [ "$lib" = default ]; then unset SDB200_LIB; else export SDB200_LIB=$lib; fi
  tag=$(basename "$lib" .so)
  python bench.py --messages $n --steps 3 --warmup 3 --no-cpu --no-lines > gpurun_out/var_$tag.json 2> gpurun_out/var_$tag.err || tail -3 gpurun_out/var_$tag.err
  python -c "
import json,sys;d=json.load(open('gpurun_out/var_$tag.json'));print('$tag','mixed',round(d['value']/1e6,2),'e2e',round(d['e2e']['value']/1e6,2),{k:round(v['msgs_per_s']/1e6,2) for k,v in d['per_kernel'].items()})"
done
unset SDB200_LIB

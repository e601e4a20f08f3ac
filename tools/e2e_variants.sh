#!/bin/bash
# usage (under gpurun): tools/e2e_variants.sh  -> e2e with the four kinds concurrent / sequential, default and variant libraries
run() { tag=$1; shift; python bench.py --messages 10000000 --steps 5 --warmup 3 --no-cpu --no-lines "$@" > gpurun_out/e2e_$tag.json 2> gpurun_out/e2e_$tag.err || tail -3 gpurun_out/e2e_$tag.err
  python -c "
import json;d=json.load(open('gpurun_out/e2e_$tag.json'));print('$tag','dev',round(d['value']/1e6,2),'e2e',round(d['e2e']['value']/1e6,2),{k:round(v,1) for k,v in d['e2e']['host_call_ms_per_kind'].items()})"; }
run conc
run seq --e2e-sequential
for lib in "$@"; do export SDB200_LIB=$lib; t=$(basename $lib .so); run ${t}_conc; run ${t}_seq --e2e-sequential; unset SDB200_LIB; done

#!/bin/bash
# usage (under gpurun): tools/ab_bench.sh <messages> [variant.so ...]   A/B of the committed baseline tree (build_variants/base_tree,
# `git archive` of the commit being compared against, built in place) against the working tree and its variants
n=$1; shift
( cd build_variants/base_tree && python bench.py --messages $n --steps 3 --warmup 3 --no-cpu --no-lines > ../../gpurun_out/var_base.json 2> ../../gpurun_out/var_base.err || tail -3 ../../gpurun_out/var_base.err
  python -c "
import json;d=json.load(open('../../gpurun_out/var_base.json'));print('BASE','mixed',round(d['value']/1e6,2),'e2e',round(d['e2e']['value']/1e6,2),{k:round(v['msgs_per_s']/1e6,2) for k,v in d['per_kernel'].items()})" )
tools/variant_bench.sh $n "$@"

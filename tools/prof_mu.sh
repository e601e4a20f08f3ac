#!/bin/bash
# usage: tools/prof_mu.sh <tag> [kernel-regex]   (run under gpurun) -> gpurun_out/prof_<tag>.ncu-rep
tag=$1; k=${2:-scan_kernel}
ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/prof_$tag -f python tools/profile_run.py MU 40000 3 > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log

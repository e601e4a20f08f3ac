#!/bin/bash
# usage: tools/prof_mu.sh <tag> [kernel-regex] [count] [messages]   (run under gpurun) -> gpurun_out/prof_<tag>.ncu-rep
tag=$1; k=${2:-scan_kernel}; c=${3:-1}; n=${4:-40000}
ncu --set full --clock-control none --import-source on -k regex:$k -s $c -c $c -o gpurun_out/prof_$tag -f python tools/profile_run.py MU $n 3 > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log

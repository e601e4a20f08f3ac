#!/bin/bash
# usage: tools/prof_kind.sh <tag> <MS|MU|MC|MN> <kernel-regex> <skip> <count> <messages>  (run under gpurun)
tag=$1; kind=$2; k=$3; s=$4; c=$5; n=$6
ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -o gpurun_out/prof_$tag -f python tools/profile_run.py $kind $n 3 > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log

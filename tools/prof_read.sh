#!/bin/bash
# usage: tools/prof_read.sh <tag> <mangled-kernel-substring> [lo-hi]   (here, after gpurun)
tag=$1; kern=$2; range=${3:-0-100000}
mkdir -p /tmp/prof && cd /tmp/prof
ncu -i /root/repo/gpurun_out/prof_$tag.ncu-rep --page source --csv > src_$tag.csv 2>/dev/null
cuobjdump -xelf all /root/repo/pysignalduino_b200/libsdb200.so > /dev/null
nvdisasm -g -c sdb_pulse.sm_100a.cubin > pulse.sass 2>/dev/null
python /root/repo/tools/ncu_summary.py /root/repo/gpurun_out/prof_$tag.ncu-rep 2>/dev/null | grep -E "duration|inst_executed.sum |issue_active|thread_inst|no_instruction|dram__bytes"
python /root/repo/tools/ncu_funcs.py src_$tag.csv pulse.sass $kern
python /root/repo/tools/ncu_lines2.py src_$tag.csv pulse.sass $kern 0 $range | sort -k4 -n -r | head -${4:-30}

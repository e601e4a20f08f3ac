#!/bin/bash
# usage (under gpurun, one B200): tools/final_measure.sh <tag>   -> everything profiles/ is refreshed from (see tools/refresh_profiles.py)
tag=${1:-r2_final}
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${tag}_pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || tail -3 gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
python tools/verify_10m.py > gpurun_out/${tag}_verify_10m.json 2> gpurun_out/${tag}_verify_10m.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --messages 1000000 --steps 2 --warmup 3 --no-cpu --no-lines > gpurun_out/${tag}_ncu_launches.log 2>&1
tools/prof_mu.sh $tag "resolve_kernel|mu_match|mu_emit|scan_kernel" 6 400000
ncu --set full --clock-control none --import-source on -k regex:format_kernel -s 2 -c 2 -o gpurun_out/prof_${tag}_format -f \
    python bench.py --messages 1000000 --steps 1 --warmup 3 --no-cpu --no-lines > gpurun_out/ncu_${tag}_format.log 2>&1
tail -2 gpurun_out/${tag}_pytest_gpu.log

#!/bin/bash
# quick GPU iteration: MU/MS parity on the corpus + fuzz, then a 2M-message bench without the CPU leg
python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -3
python bench.py --messages 2000000 --steps 3 --warmup 3 --no-cpu --no-lines > gpurun_out/quick.json 2> gpurun_out/quick.err || tail -5 gpurun_out/quick.err
python -c "
import json;d=json.load(open('gpurun_out/quick.json'));print('mixed',round(d['value']/1e6,2),'e2e',round(d['e2e']['value']/1e6,2),{k:round(v['msgs_per_s']/1e6,2) for k,v in d['per_kernel'].items()})"

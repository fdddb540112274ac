#!/bin/bash
mkdir -p gpurun_out
F="--steps 30 --warmup 5 --no-cpu-baseline --no-infer --no-other-configs --no-raw-pipeline --no-parity-check"
show() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4))
"; }
timeout 300 python bench.py $F 2>/dev/null | show "EW16"
R2N2_FWD3_EW=8 timeout 300 python bench.py $F 2>/dev/null | show "EW8"
timeout 300 python bench.py $F 2>/dev/null | show "EW16 again"
timeout 900 python -m pytest tests/test_nets_gpu.py tests/test_bench_parity_gpu.py -m gpu -x -q 2>&1 | tail -2

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_nets_gpu.py tests/test_rowpack_gpu.py tests/test_bench_parity_gpu.py tests/test_bf16x3_gpu.py tests/test_input_pipeline_gpu.py -m gpu -q -x 2>&1 | tail -3
F="--steps 30 --warmup 5 --no-cpu-baseline --no-other-configs --no-raw-pipeline --no-infer"
show() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), d['parity_check']['grad_rel_l2'], d['parity_check']['max_abs_dp'])
"; }
timeout 300 python bench.py $F --profile-out gpurun_out/r02_per_launch_e.json 2>/dev/null | show "new"
timeout 300 python bench.py $F 2>/dev/null | show "new again"

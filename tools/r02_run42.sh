#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "lrelu or pool" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_nets_gpu.py tests/test_bench_parity_gpu.py -m gpu -q -x 2>&1 | tail -3
F="--steps 40 --warmup 5 --no-cpu-baseline --no-other-configs --no-raw-pipeline --no-parity-check --no-infer"
run() { env $1 timeout 300 python bench.py $F $3 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$1', d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'])"; }
run A=1 n "--profile-out gpurun_out/r02_per_launch_n.json"
run R2N2_NO_LRELU_BIAS=1 k
run A=1 j
run R2N2_NO_LRELU_BIAS=1 k

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gru_fused_gpu.py -m gpu -q -x 2>&1 | tail -15
timeout 900 python -m pytest tests/test_nets_gpu.py tests/test_bench_parity_gpu.py -m gpu -q -x 2>&1 | tail -5
F="--steps 30 --warmup 5 --no-cpu-baseline --no-other-configs --no-raw-pipeline --no-parity-check"
run() { echo "== $1"; env $1 timeout 300 python bench.py $F --profile-out gpurun_out/r02_per_launch_$2.json 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d.get('infer'))"; }
run A=1 j
run R2N2_NO_GRU_FUSE=1 k

#!/bin/bash
# round 2, call 2: host-vs-device input diagnostic, RES_PRE + bf16x3 tests, bf16x3 bench line
mkdir -p gpurun_out
timeout 600 python tools/diag_head.py > gpurun_out/r02_diag_head.log 2>&1; echo "diag rc $?" >> gpurun_out/r02_diag_head.log
timeout 900 python -m pytest tests/test_bf16x3_gpu.py -x -q -s > gpurun_out/r02_bf16x3_tests.log 2>&1; echo "rc $?" >> gpurun_out/r02_bf16x3_tests.log
timeout 600 python bench.py --compute bf16x3 --steps 10 --warmup 3 --no-cpu-baseline --no-infer --no-raw-pipeline --no-other-configs --profile-out gpurun_out/r02_per_launch_bf16x3.json > gpurun_out/r02_bench_bf16x3.log 2> gpurun_out/r02_bench_bf16x3.err; echo "bench rc $?" >> gpurun_out/r02_bench_bf16x3.err

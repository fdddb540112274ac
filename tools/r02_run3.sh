#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_nets_gpu.py tests/test_dp_equivalence_gpu.py tests/test_bench_parity_gpu.py -m gpu -q -s > gpurun_out/r02_gputests3b.log 2>&1; echo "gpu tests rc $?" >> gpurun_out/r02_gputests3b.log
tail -8 gpurun_out/r02_gputests3b.log

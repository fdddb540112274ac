#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fwd3_gpu.py tests/test_kernels_gpu.py tests/test_up2_gpu.py -m gpu -x -q 2>&1 | tail -3
echo "default"; timeout 300 python tools/bench_conv.py --what fwd --iters 10 2>&1 | grep -v "^\[" 
echo "conv1b PAIR=1 STAGE=0"; R2N2_FWD3_STAGE_EPI=0 timeout 100 python tools/bench_conv.py --only conv1b,conv1a,conv2a --what fwd --iters 10 2>&1 | tail -3
echo "conv1b PAIR=0 STAGE=0"; R2N2_FWD3_PAIR=0 R2N2_FWD3_STAGE_EPI=0 timeout 100 python tools/bench_conv.py --only conv1b --what fwd --iters 10 2>&1 | tail -1

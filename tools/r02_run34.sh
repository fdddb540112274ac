#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_wgrad_umma_gpu.py -m gpu -q -x 2>&1 | tail -3
L=conv9b,conv10a,conv10b,conv10c,conv11
echo NSTACK; R2N2_UMMA_DBG=1 timeout 300 python tools/bench_conv.py --graph --what wgrad --iters 10 --only $L 2>&1 | grep -E "wgrad" | awk '!seen[$0]++'
echo OLD; R2N2_WGRAD3_NO_NSTACK=1 timeout 300 python tools/bench_conv.py --graph --what wgrad --iters 10 --only $L 2>&1 | grep -E "wgrad " | awk '!seen[$0]++'

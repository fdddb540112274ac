#!/bin/bash
timeout 900 python -m pytest tests/test_fwd3_gpu.py tests/test_wgrad_umma_gpu.py tests/test_gru_fused_gpu.py -m gpu -q -x 2>&1 | tail -2
OLD=$PWD/3d-r2n2_b200/libr2n2_b200_old.so
L=conv1b,conv1a,conv9b,conv10a,conv10b,conv11,conv2b,conv3b
for i in 1 2; do
echo NEW; timeout 300 python tools/bench_conv.py --graph --what dgrad,fwd --iters 10 --only $L 2>&1 | grep -E "dgrad"
echo OLD; R2N2_LIB_PATH=$OLD timeout 300 python tools/bench_conv.py --graph --what dgrad,fwd --iters 10 --only $L 2>&1 | grep -E "dgrad"
done

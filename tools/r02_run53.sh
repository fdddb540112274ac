#!/bin/bash
timeout 900 python -m pytest tests/test_fwd3_gpu.py tests/test_kernels_gpu.py tests/test_up2_gpu.py tests/test_gru_fused_gpu.py -m gpu -q -x 2>&1 | tail -2
L=conv2a,conv2b,conv3a,conv3b,conv4a,conv8a,conv9b,conv10a,conv10b,conv11,gru_wh_ur
for i in 1 2; do
echo NEW; timeout 300 python tools/bench_conv.py --graph --what fwd --iters 20 --only $L 2>&1 | grep -E " fwd "
echo OLD; R2N2_NO_SMEM_BIAS=1 timeout 300 python tools/bench_conv.py --graph --what fwd --iters 20 --only $L 2>&1 | grep -E " fwd "
done

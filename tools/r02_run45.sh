#!/bin/bash
timeout 900 python -m pytest tests/test_fwd3_gpu.py -m gpu -q -x 2>&1 | tail -2
L=conv1a,conv1b,conv2a
for i in 1 2; do
echo NEW; timeout 300 python tools/bench_conv.py --graph --what fwd --iters 20 --only $L 2>&1 | grep -E " fwd "
echo OLD; R2N2_FWD3_SKIP=64 timeout 300 python tools/bench_conv.py --graph --what fwd --iters 20 --only $L 2>&1 | grep -E " fwd "
done

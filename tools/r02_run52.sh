#!/bin/bash
timeout 900 python -m pytest tests/test_fwd3_gpu.py -m gpu -q -x 2>&1 | tail -2
for i in 1 2; do
echo NEW; timeout 300 python tools/bench_conv.py --graph --what fwd --iters 20 --only conv1a,conv1b 2>&1 | grep -E "conv1"
echo OLD; R2N2_FWD3_SKIP=64 timeout 300 python tools/bench_conv.py --graph --what fwd --iters 20 --only conv1a,conv1b 2>&1 | grep -E "conv1"
done

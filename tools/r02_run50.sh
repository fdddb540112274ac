#!/bin/bash
for s in 43 59 27; do
echo "SKIP=$s"; R2N2_FWD3_SKIP=$s timeout 300 python tools/bench_conv.py --graph --what dgrad,fwd --iters 10 --only conv1b,conv1a 2>&1 | grep -E "conv1"
done
echo "EW=8 full"; R2N2_FWD3_EW=8 timeout 300 python tools/bench_conv.py --graph --what dgrad,fwd --iters 10 --only conv1b,conv1a 2>&1 | grep -E "conv1"
echo "EW=8 SKIP=43"; R2N2_FWD3_EW=8 R2N2_FWD3_SKIP=43 timeout 300 python tools/bench_conv.py --graph --what dgrad,fwd --iters 10 --only conv1b,conv1a 2>&1 | grep -E "conv1"

#!/bin/bash
timeout 900 python -m pytest tests/test_fwd3_gpu.py -m gpu -x -q 2>&1 | tail -3
echo "EW16"; timeout 200 python tools/bench_conv.py --only conv1a,conv1b --what fwd --iters 10 2>&1 | tail -2
echo "EW8"; R2N2_FWD3_EW=8 timeout 200 python tools/bench_conv.py --only conv1a,conv1b --what fwd --iters 10 2>&1 | tail -2
for s in 9 4; do echo "EW16 SKIP=$s"; R2N2_FWD3_SKIP=$s timeout 100 python tools/bench_conv.py --only conv1a,conv1b --what fwd --iters 10 2>&1 | tail -2; done

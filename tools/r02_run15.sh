#!/bin/bash
for s in 13 29 61; do echo "pair SKIP=$s"; R2N2_FWD3_SKIP=$s timeout 100 python tools/bench_conv.py --only conv1b --what fwd --iters 10 2>&1 | tail -1; done

#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r02_fwd3_skip.txt
for s in 0 1 2 4 8 3 5 6 7 12 14 15; do
  echo "== R2N2_FWD3_SKIP=$s (1 planes, 2 weights, 4 epilogue, 8 MMA switched off)" >> gpurun_out/r02_fwd3_skip.txt
  R2N2_FWD3_SKIP=$s timeout 120 python tools/bench_conv.py --only conv1b,conv9b,conv10b,conv11 --what fwd --iters 10 >> gpurun_out/r02_fwd3_skip.txt 2>&1
done
cat gpurun_out/r02_fwd3_skip.txt

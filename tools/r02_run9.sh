#!/bin/bash
for s in 0 8 9 12 24 40 56 25 57 4 5 16 32 48; do echo "SKIP=$s (1 planes 2 weights 4 epilogue 8 MMA 16 tmem-ld 32 st.global)"; R2N2_FWD3_SKIP=$s timeout 120 python tools/bench_conv.py --only conv1b --what fwd --iters 10 2>&1 | tail -1; done

#!/bin/bash
for s in 0 9 57 4; do echo "SKIP=$s (1 planes 2 weights 4 epilogue 8 MMA 16 tmem-ld 32 st.global)"; R2N2_FWD3_SKIP=$s timeout 120 python tools/bench_conv.py --only conv1b --what fwd --iters 10 2>&1 | tail -1; done
echo PAIR=0; R2N2_FWD3_PAIR=0 timeout 120 python tools/bench_conv.py --only conv1b --what fwd --iters 10 2>&1 | tail -1
echo PAIR=0 TMA_EPI=1; R2N2_FWD3_PAIR=0 R2N2_FWD3_TMA_EPI=1 timeout 120 python tools/bench_conv.py --only conv1b --what fwd --iters 10 2>&1 | tail -1
echo PAIR=1 TMA_EPI=1; R2N2_FWD3_PAIR=1 R2N2_FWD3_TMA_EPI=1 timeout 120 python tools/bench_conv.py --only conv1b --what fwd --iters 10 2>&1 | tail -1
timeout 600 python -m pytest tests/test_fwd3_gpu.py -m gpu -x -q 2>&1 | tail -3

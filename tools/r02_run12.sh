#!/bin/bash
echo default; timeout 100 python tools/bench_conv.py --only conv1a --what fwd --iters 10 2>&1 | tail -1
echo K71 PAIR=0; R2N2_UMMA_DBG=1 R2N2_FWD3_K71=1 R2N2_FWD3_PAIR=0 timeout 100 python tools/bench_conv.py --only conv1a --what fwd --iters 10 2>&1 | grep -E "fwd3\]|conv" | sort | uniq -c
echo K71 PAIR=1; R2N2_UMMA_DBG=1 R2N2_FWD3_K71=1 R2N2_FWD3_PAIR=1 timeout 100 python tools/bench_conv.py --only conv1a --what fwd --iters 10 2>&1 | grep -E "fwd3\]|conv" | sort | uniq -c
echo K71 PAIR=0 STAGE=0; R2N2_FWD3_K71=1 R2N2_FWD3_PAIR=0 R2N2_FWD3_STAGE_EPI=0 timeout 100 python tools/bench_conv.py --only conv1a --what fwd --iters 10 2>&1 | tail -1
for s in 4 9 12; do echo "K71 PAIR=1 SKIP=$s"; R2N2_FWD3_SKIP=$s R2N2_FWD3_K71=1 R2N2_FWD3_PAIR=1 timeout 100 python tools/bench_conv.py --only conv1a --what fwd --iters 10 2>&1 | tail -1; done

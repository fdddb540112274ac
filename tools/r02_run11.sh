#!/bin/bash
echo "default"; timeout 200 python tools/bench_conv.py --only conv2a,conv2a_dgrad,conv2b --what fwd --iters 10 2>&1 | tail -3
echo "PAIR=0 FORCE=1"; R2N2_UMMA_DBG=1 R2N2_FWD3_PAIR=0 R2N2_FWD3_FORCE=1 timeout 200 python tools/bench_conv.py --only conv2a,conv2a_dgrad,conv2b --what fwd --iters 10 2>&1 | grep -E "fwd3\]|conv" | sort | uniq -c
echo "PAIR=0 FORCE=1 STAGE_EPI=0"; R2N2_FWD3_STAGE_EPI=0 R2N2_FWD3_PAIR=0 R2N2_FWD3_FORCE=1 timeout 200 python tools/bench_conv.py --only conv2a,conv2a_dgrad,conv2b --what fwd --iters 10 2>&1 | tail -3
echo "PAIR=0 FORCE=1 WT=3"; R2N2_FWD3_WT=3 R2N2_FWD3_PAIR=0 R2N2_FWD3_FORCE=1 timeout 200 python tools/bench_conv.py --only conv2a,conv2a_dgrad,conv2b --what fwd --iters 10 2>&1 | tail -3
echo "PAIR=1 FORCE=1 conv2a_dgrad"; R2N2_UMMA_DBG=1 R2N2_FWD3_PAIR=1 R2N2_FWD3_FORCE=1 timeout 200 python tools/bench_conv.py --only conv2a_dgrad --what fwd --iters 10 2>&1 | grep -E "fwd3\]|conv" | sort | uniq -c

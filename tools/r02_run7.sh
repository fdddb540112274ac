#!/bin/bash
mkdir -p gpurun_out
R2N2_UMMA_DBG=1 timeout 120 python tools/bench_conv.py --only conv1b --what fwd --iters 10 2>&1 | grep -E "fwd3\]|conv1b" | sort | uniq -c
R2N2_FWD3_PAIR=0 timeout 120 python tools/bench_conv.py --only conv1b --what fwd --iters 10 2>&1 | tail -1
for s in 1 4 8 12 5 ; do echo "skip $s"; R2N2_FWD3_SKIP=$s timeout 120 python tools/bench_conv.py --only conv1b --what fwd --iters 10 2>&1 | tail -1; done
R2N2_FWD3_PAIR=1 R2N2_FWD3_FORCE=1 R2N2_UMMA_DBG=1 timeout 120 python tools/bench_conv.py --only conv2a,conv2b,conv9b,conv10a,conv10b --what fwd --iters 10 2>&1 | grep -E "fwd3\]|conv" | sort | uniq -c
R2N2_WAITSTAT=1 timeout 300 python tools/waitstat.py --only conv1b --what fwd,dgrad 2>&1 | grep -v "^\[umma"

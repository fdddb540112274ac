#!/bin/bash
L=conv9b,conv10a,conv10b,conv11
b() { echo "== $1"; env $1 R2N2_UMMA_DBG=1 timeout 300 python tools/bench_conv.py --graph --what wgrad --iters 10 --only $L 2>&1 | grep -E "wgrad" | awk '!seen[$0]++' | sed 's/\[umma wgrad3\] //' | cut -c1-150; }
b A=1
b R2N2_WGRAD3_Y=2
b R2N2_WGRAD3_Y=4
b R2N2_WGRAD3_Y=8
b R2N2_WGRAD3_Y=16
b R2N2_WGRAD3_ONE_CTA=1
b R2N2_WGRAD3_SPLITS=148
b R2N2_WGRAD3_SPLITS=74
b R2N2_WGRAD3_STAGES=2
b R2N2_WGRAD3_STAGES=3

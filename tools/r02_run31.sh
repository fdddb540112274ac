#!/bin/bash
R2N2_UMMA_DBG=1 timeout 600 python tools/bench_conv.py --what wgrad --iters 10 2>&1 | grep -E "wgrad" | awk '!seen[$0]++'

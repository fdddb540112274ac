#!/bin/bash
echo default; timeout 300 python tools/bench_conv.py --graph --what wgrad --iters 10 --only conv1b,conv2a,conv2a_dgrad 2>&1 | grep -E "wgrad" 
echo CHUNK64; R2N2_UMMA_DBG=1 R2N2_WGRAD_CHUNK64=1 timeout 300 python tools/bench_conv.py --graph --what wgrad --iters 10 --only conv1b,conv2a,conv2a_dgrad 2>&1 | grep -E "wgrad" | awk '!seen[$0]++'

#!/bin/bash
L=conv4a,conv5a,conv6a,fc7,xproj_ur,xproj_ur_dgrad,gru_wh_ur,conv7a
echo "new"; timeout 300 python tools/bench_conv.py --graph --only $L --what fwd --iters 20 2>&1 | grep -E "^conv|^fc|^xp|^gru"
echo "old"; R2N2_UMMA_NO_MC=1 timeout 300 python tools/bench_conv.py --graph --only $L --what fwd --iters 20 2>&1 | grep -E "^conv|^fc|^xp|^gru"
echo "old + 2 ctas budget"; R2N2_UMMA_NO_MC=1 R2N2_UMMA_CTAS=2 timeout 300 python tools/bench_conv.py --graph --only $L --what fwd --iters 20 2>&1 | grep -E "^conv|^fc|^xp|^gru"

#!/bin/bash
L=conv9b,conv10a,conv10a_dgrad,conv10b,conv11,conv11_dgrad
echo default; timeout 200 python tools/bench_conv.py --only $L --what fwd --iters 10 2>&1 | tail -6
echo STK=1; R2N2_FWD3_STK=1 timeout 200 python tools/bench_conv.py --only $L --what fwd --iters 10 2>&1 | tail -6
echo PAIR=1; R2N2_FWD3_PAIR=1 timeout 200 python tools/bench_conv.py --only $L --what fwd --iters 10 2>&1 | tail -6

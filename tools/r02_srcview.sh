#!/bin/bash
# ncu source-level captures (stall samples per SASS instruction) of single launches: tools/bench_conv.py under ncu
mkdir -p gpurun_out
cap() { # name, what, layer, kernel regex
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"$4" --launch-skip 1 -c 1 -o gpurun_out/sv_$1 python tools/bench_conv.py --what $2 --iters 1 --only $3 > gpurun_out/ncu_sv_$1.log 2>&1
echo "$1 rc $?"
ncu -i gpurun_out/sv_$1.ncu-rep --page source --csv > gpurun_out/sv_$1_source.csv 2>/dev/null
rm -f gpurun_out/sv_$1.ncu-rep; }
cap conv2b_fwd fwd conv2b conv_fwd_umma_persistent
cap conv2b_dgrad dgrad conv2b conv_fwd_umma_persistent
cap conv9b_dgrad dgrad conv9b conv_fwd3_umma_kernel
cap conv3a_wgrad wgrad conv3a conv_wgrad_umma_kernel

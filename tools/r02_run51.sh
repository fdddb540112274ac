#!/bin/bash
mkdir -p gpurun_out
cap() { # name, what, layer
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"$4" --launch-skip 1 -c 1 -o gpurun_out/r02d_$1_src python tools/bench_conv.py --what $2 --iters 1 --only $3 > gpurun_out/ncu_$1.log 2>&1
echo "$1 rc $?"
ncu -i gpurun_out/r02d_$1_src.ncu-rep --page source --csv > gpurun_out/r02d_$1_source.csv 2>/dev/null
rm -f gpurun_out/r02d_$1_src.ncu-rep; }
cap conv1b_dgrad dgrad conv1b conv_fwd3_umma_kernel
cap conv1b_fwd fwd conv1b conv_fwd3_umma_kernel

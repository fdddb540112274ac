#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_fwd3_umma_kernel --launch-skip 1 -c 1 -o gpurun_out/r02c_conv1a_src python tools/bench_conv.py --what fwd --iters 1 --only conv1a > gpurun_out/ncu_conv1a.log 2>&1
echo rc $?
ncu -i gpurun_out/r02c_conv1a_src.ncu-rep --page source --csv > gpurun_out/r02c_conv1a_source.csv 2>/dev/null
ncu -i gpurun_out/r02c_conv1a_src.ncu-rep --page details --csv > gpurun_out/r02c_conv1a_details.csv 2>/dev/null
rm -f gpurun_out/r02c_conv1a_src.ncu-rep
ls -la gpurun_out/r02c_conv1a*

#!/bin/bash
# Round profile evidence (run under gpurun): plain bench, ncu launch list of the same command, and
# `ncu --set full` captures of the dominant kernels.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
timeout 600 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_full.log
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/ncu_launches.log
# persistent per-tap kernel: launch 0 = conv1a (TMA epilogue), launch 3 = conv2b (N=128, K=1152)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_umma_persistent -s 3 -c 1 \
    -o gpurun_out/ncu_conv2b_fwd $CMD > gpurun_out/ncu_full1.log 2>&1
echo "ncu full1 exit $?" >> gpurun_out/ncu_full1.log
# flat-halo forward: launches 0-4 = conv1b (one per view), 5 = conv9b (dx-stacked), 6 = conv10a (dx-stacked), 7 = conv10b
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_fwd3_umma -s 5 -c 1 \
    -o gpurun_out/ncu_conv9b_fwd3_stk $CMD > gpurun_out/ncu_full2.log 2>&1
echo "ncu full2 exit $?" >> gpurun_out/ncu_full2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_umma_kernel -s 30 -c 1 \
    -o gpurun_out/ncu_wgrad_v1 $CMD > gpurun_out/ncu_full3.log 2>&1
echo "ncu full3 exit $?" >> gpurun_out/ncu_full3.log

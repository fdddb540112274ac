#!/bin/bash
# Round profile evidence (run under gpurun): plain bench, ncu launch list of the same command, and one
# `ncu --set full` capture of the dominant kernels.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
timeout 600 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_full.log
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/ncu_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_umma_persistent -s 1 -c 1 \
    -o gpurun_out/ncu_conv1b_fwd $CMD > gpurun_out/ncu_full1.log 2>&1
echo "ncu full1 exit $?" >> gpurun_out/ncu_full1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad3_umma -s 2 -c 1 \
    -o gpurun_out/ncu_wgrad3 $CMD > gpurun_out/ncu_full2.log 2>&1
echo "ncu full2 exit $?" >> gpurun_out/ncu_full2.log

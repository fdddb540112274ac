#!/bin/bash
# round-2 profile evidence (one GPU): gate micro-benchmark, per-step DRAM traffic (launch list), ncu --set full of the
# conv kernels of ONE step, summaries written next to the reports
mkdir -p gpurun_out
timeout 300 python tools/bench_gates.py --json gpurun_out/r02_gates_b2304.json > gpurun_out/r02_gates_b2304.txt 2>&1
timeout 300 python tools/bench_gates.py --batch 36 --json gpurun_out/r02_gates_b36.json > gpurun_out/r02_gates_b36.txt 2>&1
timeout 300 python tools/one_step.py --calls-out gpurun_out/r02b_calls.json > gpurun_out/one_step_plain.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02b_step_traffic.csv python tools/one_step.py > gpurun_out/ncu_traffic.log 2>&1
echo "ncu traffic rc $?" >> gpurun_out/ncu_traffic.log
timeout 1200 ncu --profile-from-start off --set full --clock-control none -k regex:'conv_fwd3_umma_kernel' -c 16 -o gpurun_out/r02_ncu_fwd3 python tools/one_step.py > gpurun_out/ncu_fwd3.log 2>&1
echo "ncu fwd3 rc $?" >> gpurun_out/ncu_fwd3.log
timeout 1200 ncu --profile-from-start off --set full --clock-control none -k regex:'conv_fwd_umma_persistent' -c 12 -o gpurun_out/r02_ncu_fwd python tools/one_step.py > gpurun_out/ncu_fwd.log 2>&1
echo "ncu fwd rc $?" >> gpurun_out/ncu_fwd.log
timeout 1200 ncu --profile-from-start off --set full --clock-control none -k regex:'conv_wgrad' -c 31 -o gpurun_out/r02_ncu_wgrad python tools/one_step.py > gpurun_out/ncu_wgrad.log 2>&1
echo "ncu wgrad rc $?" >> gpurun_out/ncu_wgrad.log
for r in r02_ncu_fwd3 r02_ncu_fwd r02_ncu_wgrad; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
  rm -f gpurun_out/$r.ncu-rep
done
ls -la gpurun_out | tail -12
cat gpurun_out/r02_gates_b2304.txt gpurun_out/r02_gates_b36.txt

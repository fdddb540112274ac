#!/bin/bash
# round-2 (session 3) profile evidence, one GPU: full GPU test suite, per-step DRAM traffic (launch list), ncu --set full of
# the conv kernels of ONE step (forward / data gradient, fused-GRU variants, weight gradients) and of the bandwidth-bound
# kernels that changed (pool code kernels, LeakyReLU' + column sums); raw pages as csv, reports dropped
mkdir -p gpurun_out
if [ -z "$R2N2_PROFILE_SKIP_TESTS" ]; then
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02c_gputests.log
cat gpurun_out/r02c_gputests.log
fi
timeout 300 python tools/one_step.py --calls-out gpurun_out/r02c_calls.json > gpurun_out/one_step_plain.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02c_step_traffic.csv python tools/one_step.py > gpurun_out/ncu_traffic.log 2>&1
echo "ncu traffic rc $?"
N="ncu --profile-from-start off --set full --clock-control none"
timeout 1200 $N -k regex:'conv_fwd3_umma_kernel' -c 16 -o gpurun_out/r02c_ncu_fwd3 python tools/one_step.py > gpurun_out/ncu_fwd3.log 2>&1; echo "fwd3 rc $?"
timeout 1200 $N -k regex:'conv_fwd_umma_persistent' -c 12 -o gpurun_out/r02c_ncu_fwd python tools/one_step.py > gpurun_out/ncu_fwd.log 2>&1; echo "fwd rc $?"
# (persistent-kernel launches 27-30 of a step)
timeout 1200 $N -k regex:'conv_fwd_umma_persistent' --launch-skip 26 -c 4 -o gpurun_out/r02c_ncu_gru python tools/one_step.py > gpurun_out/ncu_gru.log 2>&1; echo "gru rc $?"
timeout 1200 $N -k regex:'conv_wgrad' -c 31 -o gpurun_out/r02c_ncu_wgrad python tools/one_step.py > gpurun_out/ncu_wgrad.log 2>&1; echo "wgrad rc $?"
timeout 1200 $N -k regex:'maxpool_|lrelu_bwd_colsum' -c 16 -o gpurun_out/r02c_ncu_bw python tools/one_step.py > gpurun_out/ncu_bw.log 2>&1; echo "bw rc $?"
for r in r02c_ncu_fwd3 r02c_ncu_fwd r02c_ncu_gru r02c_ncu_wgrad r02c_ncu_bw; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
  rm -f gpurun_out/$r.ncu-rep
done
ls -la gpurun_out | tail -12

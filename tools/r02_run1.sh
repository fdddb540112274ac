#!/bin/bash
# round 2, call 1: new parity tests (numbers), full GPU suite, bench, per-step traffic + bandwidth-kernel ncu
mkdir -p gpurun_out
export R2N2_PARITY_OUT=gpurun_out/r02_parity.jsonl
rm -f $R2N2_PARITY_OUT
timeout 900 python -m pytest tests/test_bench_parity_gpu.py -q -s > gpurun_out/r02_parity_tests.log 2>&1; echo "parity rc $?" >> gpurun_out/r02_parity_tests.log
unset R2N2_PARITY_OUT
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_bench_parity_gpu.py > gpurun_out/r02_gputests.log 2>&1; echo "gpu tests rc $?" >> gpurun_out/r02_gputests.log
timeout 600 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/r02_per_launch_a.json > gpurun_out/r02_bench_a.log 2> gpurun_out/r02_bench_a.err; echo "bench rc $?" >> gpurun_out/r02_bench_a.err
timeout 300 python tools/one_step.py --calls-out gpurun_out/r02_calls.json > gpurun_out/one_step_plain.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_step_traffic.csv python tools/one_step.py > gpurun_out/ncu_traffic.log 2>&1
echo "ncu traffic rc $?" >> gpurun_out/ncu_traffic.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'adam_kernel|softmax_ce3d_kernel|maxpool_bwd_kernel|maxpool_fwd_kernel|gru_ur_fwd_kernel|gru_out_bwd_kernel|lrelu_bwd_kernel|bias_grad_flat_kernel|unpool_fwd_kernel' -o gpurun_out/r02_ncu_bw python tools/one_step.py > gpurun_out/ncu_bw.log 2>&1
echo "ncu bw rc $?" >> gpurun_out/ncu_bw.log

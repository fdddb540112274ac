#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02c_bench_full.log 2> gpurun_out/r02c_bench_full.err; echo "default rc $?"
grep '^{' gpurun_out/r02c_bench_full.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['infer'], d['other_configs']['lstm_train']['value'], d['other_configs']['infer_24view'], d['parity_check'], d['cpu_baseline']['value'])"
R2N2_COMPUTE=bf16x3 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs --no-raw-pipeline --no-infer > gpurun_out/r02c_bench_bf16x3.log 2>/dev/null; echo "bf16x3 rc $?"
grep '^{' gpurun_out/r02c_bench_bf16x3.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'], d['parity_check'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02c_bench_ref.log 2>/dev/null; echo "ref rc $?"; tail -c 600 gpurun_out/r02c_bench_ref.log

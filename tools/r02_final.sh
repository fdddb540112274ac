#!/bin/bash
# end-of-round validation: full GPU suite, smoke(), default bench line (as the driver runs it), bf16x3 line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r02e_gputests.log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -8 | tee gpurun_out/r02e_smoke.log
timeout 900 python bench.py > gpurun_out/r02e_bench_full.log 2> gpurun_out/r02e_bench_full.err; echo "default rc $?"
grep '^{' gpurun_out/r02e_bench_full.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['infer'], d['other_configs']['lstm_train']['value'], d['other_configs']['infer_24view']['ms_per_batch'], d['parity_check']['max_abs_dp'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])"
R2N2_COMPUTE=bf16x3 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs --no-raw-pipeline --no-infer > gpurun_out/r02e_bench_bf16x3.log 2>/dev/null; echo "bf16x3 rc $?"
grep '^{' gpurun_out/r02e_bench_bf16x3.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'], d['parity_check']['max_abs_dp'])"
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-other-configs --no-raw-pipeline --no-parity-check --no-infer --profile-out gpurun_out/r02e_per_launch.json > /dev/null 2>&1; echo "profile rc $?"

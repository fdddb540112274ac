#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_nets_gpu.py -m gpu -q -x 2>&1 | tail -3
F="--steps 30 --warmup 5 --no-cpu-baseline --no-other-configs --no-raw-pipeline --no-parity-check --no-infer"
timeout 300 python bench.py $F --profile-out gpurun_out/r02_per_launch_o.json 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'])"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_per_launch_o.json'))
print(d['sum_ms'])
for k,v in sorted(d['per_kernel'].items(), key=lambda kv:-kv[1]['ms']):
    if v['ms'] < 0.12: print(k, round(v['ms'],4), v['calls'], round(v['gbs']))
PY

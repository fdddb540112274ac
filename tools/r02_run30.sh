#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fwd3_gpu.py -m gpu -q -x 2>&1 | tail -2
F="--steps 30 --warmup 5 --no-cpu-baseline --no-other-configs --no-raw-pipeline --no-infer --no-parity-check"
timeout 300 python bench.py $F --profile-out gpurun_out/r02_per_launch_f.json > /dev/null 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_per_launch_f.json'))
for k,shape,ms,extra in d['launches']:
    if ('M2903220' in shape or 'K2903220' in shape): print(k,shape,round(ms,4))
print(d['sum_ms'])
PY

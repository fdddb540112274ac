#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-infer --no-other-configs --no-raw-pipeline --profile-out gpurun_out/r02_per_launch_b.json > gpurun_out/r02_bench_b.log 2> gpurun_out/r02_bench_b.err; echo "bench rc $?"
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_b.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['parity_check'])
d=json.load(open('gpurun_out/r02_per_launch_b.json'))
for k,shape,ms,extra in d['launches']:
    if 'conv' in k and ms>0.1: print(k,shape,round(ms,4))
PY

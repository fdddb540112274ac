#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/r02_bench_full.log 2> gpurun_out/r02_bench_full.err ) 2>&1 | grep real
timeout 600 python bench.py --compute bf16x3 --steps 10 --warmup 3 --no-cpu-baseline --no-infer --no-raw-pipeline --no-other-configs > gpurun_out/r02_bench_bf16x3_b.log 2>/dev/null
( time timeout 600 python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/r02_bench_ref.log 2>/dev/null ) 2>&1 | grep real
python - <<'PY'
import json
for fn in ('gpurun_out/r02_bench_full.log','gpurun_out/r02_bench_bf16x3_b.log','gpurun_out/r02_bench_ref.log'):
    for l in open(fn):
        if l.startswith('{'):
            d=json.loads(l)
            print(fn, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', d['e2e']['value'], 'frac', (d.get('roofline') or {}).get('frac'))
            for k in ('infer','other_configs','parity_check','cpu_baseline','clocks','e2e_raw_pipeline'):
                if d.get(k): print('   ',k, json.dumps(d[k])[:600])
PY

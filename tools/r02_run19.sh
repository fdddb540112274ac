#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gputests5.log 2>&1; echo "gpu tests rc $?" >> gpurun_out/r02_gputests5.log
tail -4 gpurun_out/r02_gputests5.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-infer --no-other-configs --no-raw-pipeline --profile-out gpurun_out/r02_per_launch_d.json > gpurun_out/r02_bench_d.log 2> gpurun_out/r02_bench_d.err; echo "bench rc $?"
R2N2_FWD3_PAIR=0 R2N2_FWD3_STAGE_EPI=0 R2N2_FWD3_K71=0 R2N2_FWD3_C96=0 R2N2_FWD3_TMA_EPI=1 R2N2_UMMA_CTAS=2 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-infer --no-other-configs --no-raw-pipeline --no-parity-check > gpurun_out/r02_bench_d_old.log 2>/dev/null
python - <<'PY'
import json
for fn in ('gpurun_out/r02_bench_d.log','gpurun_out/r02_bench_d_old.log'):
    for l in open(fn):
        if l.startswith('{'):
            d=json.loads(l); print(fn, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
PY

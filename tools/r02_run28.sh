#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gputests6.log 2>&1; echo "gpu tests rc $?" >> gpurun_out/r02_gputests6.log
tail -4 gpurun_out/r02_gputests6.log
F="--steps 30 --warmup 5 --no-cpu-baseline --no-other-configs --no-raw-pipeline --no-parity-check"
show() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), d['infer'])
"; }
timeout 300 python bench.py $F 2>/dev/null | show "PDL on"
R2N2_PDL=0 timeout 300 python bench.py $F 2>/dev/null | show "PDL off"
timeout 300 python bench.py $F 2>/dev/null | show "PDL on again"

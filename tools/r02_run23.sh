#!/bin/bash
F="--steps 30 --warmup 5 --no-cpu-baseline --no-infer --no-other-configs --no-parity-check --no-raw-pipeline"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 $F 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['dp_check'])
"

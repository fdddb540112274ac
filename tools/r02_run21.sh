#!/bin/bash
mkdir -p gpurun_out
F="--steps 30 --warmup 5 --no-cpu-baseline --no-infer --no-other-configs --no-parity-check --no-raw-pipeline"
run() { tag=$1; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29560+RANDOM%100)) bench.py --gpus 2 $F 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$tag', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))
"; }
timeout 300 python bench.py --gpus 1 $F 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('n1', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))
"
R2N2_SM_LIMIT=140 timeout 300 python bench.py --gpus 1 $F 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('n1 SM_LIMIT=140', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))
"
run "n2 default(reserve 8)" A=1
run "n2 reserve 0" R2N2_COMM_SMS=0
run "n2 reserve 4" R2N2_COMM_SMS=4
run "n2 reserve 16" R2N2_COMM_SMS=16
run "n2 maxctas8 only" R2N2_COMM_SMS=0 NCCL_MAX_CTAS=8

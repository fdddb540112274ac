#!/bin/bash
mkdir -p gpurun_out
F="--steps 30 --warmup 5 --no-cpu-baseline --no-infer --no-other-configs --no-parity-check --no-raw-pipeline"
show() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d.get('dp_check',{}) and d['dp_check'].get('params_bit_identical_across_ranks'))
"; }
timeout 300 python bench.py --gpus 1 $F 2>/dev/null | show n1
run() { tag=$1; n=$2; shift; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29560+RANDOM%100)) bench.py --gpus $n $F 2>gpurun_out/r02_run22_$n.err | show "$tag"; }
run "n8 reserve 0 (round-1 behaviour)" 8 R2N2_COMM_SMS=0
run "n8 reserve 8" 8 R2N2_COMM_SMS=8
run "n8 reserve 4" 8 R2N2_COMM_SMS=4
run "n8 reserve 16" 8 R2N2_COMM_SMS=16

#!/bin/bash
mkdir -p gpurun_out
F="--gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-infer --no-parity-check"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py $F > gpurun_out/r02_n2.log 2>gpurun_out/r02_n2.err
grep '^{' gpurun_out/r02_n2.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'raw', d.get('e2e_raw_pipeline',{}).get('value'), d.get('dp_check'), d.get('other_configs'))"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-infer --no-parity-check --no-other-configs --no-raw-pipeline 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('N=1', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'])"

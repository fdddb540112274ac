#!/bin/bash
mkdir -p gpurun_out
F="--steps 30 --warmup 5 --no-cpu-baseline --no-infer --no-other-configs --no-parity-check --no-raw-pipeline"
timeout 300 python bench.py --gpus 1 $F > gpurun_out/r02_scale_n1.log 2> gpurun_out/r02_scale_n1.err
for n in 2 4 8; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540+n)) bench.py --gpus $n $F > gpurun_out/r02_scale_n$n.log 2> gpurun_out/r02_scale_n$n.err
done
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    for l in open('gpurun_out/r02_scale_n%d.log'%n):
        if l.startswith('{'):
            d=json.loads(l)
            if n==1: base=d
            print(n, round(d['value'],1), round(d['ms_per_step'],3), 'eff %.4f'%(d['value']/(n*base['value'])), 'e2e', round(d['e2e']['value'],1), 'e2e eff %.4f'%(d['e2e']['value']/(n*base['e2e']['value'])), d.get('dp_check'))
PY

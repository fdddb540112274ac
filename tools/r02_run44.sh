#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -12 > gpurun_out/r02c_topo8.txt
F="--steps 30 --warmup 5 --no-cpu-baseline --no-infer --no-parity-check --no-other-configs"
run() { env $1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 2951$3 bench.py --gpus $2 $F > gpurun_out/r02c_scale_$3.log 2>gpurun_out/r02c_scale_$3.err
grep '^{' gpurun_out/r02c_scale_$3.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$1 N=$2', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'raw', d.get('e2e_raw_pipeline',{}).get('value'), (d.get('dp_check') or {}).get('per_rank_ms_without_sync'), (d.get('dp_check') or {}).get('sync_overhead_ms'))"; }
run A=1 8 1
run R2N2_NO_NUMA_BIND=1 8 2
timeout 300 python bench.py $F > gpurun_out/r02c_scale_n1.log 2>/dev/null
grep '^{' gpurun_out/r02c_scale_n1.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('N=1', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'raw', d.get('e2e_raw_pipeline',{}).get('value'))"
run A=1 4 4
cat gpurun_out/r02c_topo8.txt | cut -c1-150

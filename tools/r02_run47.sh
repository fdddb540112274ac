#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nets_gpu.py -m gpu -q -x -k "ranges or adam" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_dp_equivalence_gpu.py -m gpu -q -x 2>&1 | tail -3
F="--gpus 2 --steps 30 --warmup 5 --no-cpu-baseline --no-infer --no-parity-check --no-other-configs --no-raw-pipeline"
run() { env $1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2950$2 bench.py $F > gpurun_out/r02d_n2_$2.log 2>gpurun_out/r02d_n2_$2.err
grep '^{' gpurun_out/r02d_n2_$2.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); c=d.get('dp_check') or {}; print('$1', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], c.get('params_bit_identical_across_ranks'), c.get('per_rank_ms_without_sync'), c.get('sync_overhead_ms'))"; }
run A=1 1
run R2N2_NO_SPLIT_UPDATE=1 2
run A=1 3
run R2N2_NO_SPLIT_UPDATE=1 4

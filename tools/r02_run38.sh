#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
F="--steps 40 --warmup 5 --no-cpu-baseline --no-other-configs --no-raw-pipeline --no-parity-check --no-infer"
run() { env $1 timeout 300 python bench.py $F $3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$1', d['ms_per_step'], d['roofline']['frac'])"; }
run A=1 l "--profile-out gpurun_out/r02_per_launch_l.json"
run R2N2_WGRAD_NO_PLAIN_STORE=1 k
run A=1 j
run R2N2_WGRAD_NO_PLAIN_STORE=1 k

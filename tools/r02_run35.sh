#!/bin/bash
mkdir -p gpurun_out
F="--steps 30 --warmup 5 --no-cpu-baseline --no-other-configs --no-raw-pipeline --no-infer --no-parity-check"
run() { echo "== $1"; env $1 timeout 300 python bench.py $F --profile-out gpurun_out/r02_per_launch_$2.json 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['clocks'])"; }
run A=1 g
run R2N2_NO_COLSUM_FUSE=1 h
run R2N2_TMA_EPI=0 i

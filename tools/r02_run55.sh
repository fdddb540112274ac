#!/bin/bash
OLD=$PWD/3d-r2n2_b200/libr2n2_b200_old.so
for i in 1 2; do
echo NEW; timeout 300 python tools/bench_conv.py --graph --what dgrad,fwd --iters 10 --only conv1b 2>&1 | grep -E "dgrad"
echo OLD; R2N2_LIB_PATH=$OLD timeout 300 python tools/bench_conv.py --graph --what dgrad,fwd --iters 10 --only conv1b 2>&1 | grep -E "dgrad"
done

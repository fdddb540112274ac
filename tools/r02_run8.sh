#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fwd3_gpu.py -m gpu -x -q > gpurun_out/r02_fwd3_tests.log 2>&1; echo "rc $?" >> gpurun_out/r02_fwd3_tests.log
tail -15 gpurun_out/r02_fwd3_tests.log
for pr in 1 0; do for se in 1 0; do echo "PAIR=$pr STAGE_EPI=$se"; R2N2_FWD3_PAIR=$pr R2N2_FWD3_STAGE_EPI=$se timeout 120 python tools/bench_conv.py --only conv1b --what fwd --iters 10 2>&1 | tail -1; done; done
echo "default"; timeout 200 python tools/bench_conv.py --only conv1b,conv9b,conv10a,conv10a_dgrad,conv10b,conv11,conv11_dgrad --what fwd --iters 10 2>&1 | tail -7
echo "STAGE_EPI=0"; R2N2_FWD3_STAGE_EPI=0 timeout 200 python tools/bench_conv.py --only conv9b,conv10a,conv10a_dgrad,conv10b,conv11,conv11_dgrad --what fwd --iters 10 2>&1 | tail -6
echo "STK=0"; R2N2_FWD3_STK=0 timeout 200 python tools/bench_conv.py --only conv9b,conv10a,conv10a_dgrad,conv10b,conv11,conv11_dgrad --what fwd --iters 10 2>&1 | tail -6

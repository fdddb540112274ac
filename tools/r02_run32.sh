#!/bin/bash
timeout 900 python -m pytest tests/test_wgrad_umma_gpu.py tests/test_up2_gpu.py tests/test_rowpack_gpu.py -m gpu -q -x 2>&1 | tail -2
echo new; timeout 600 python tools/bench_conv.py --graph --what wgrad --iters 10 2>&1 | grep -E "wgrad" 
echo old; R2N2_UMMA_WGRAD_GENERIC=1 timeout 600 python tools/bench_conv.py --graph --what wgrad --iters 10 --only conv1a,conv1b,conv2a,conv2b,conv2c,conv3a,conv3b,conv4a,conv5a,gru_wh_ur,conv7a,conv8a,conv9a,conv9a_dgrad,fc7,xproj_ur 2>&1 | grep -E "wgrad"

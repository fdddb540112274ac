#!/bin/bash
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -2
R2N2_WAITSTAT=1 timeout 300 python tools/waitstat.py --only fc7,gru_wh_ur,conv6a,conv5a,xproj_ur_dgrad,xproj_ur --what fwd 2>&1 | grep -E "==|umma fwd\]|warp  [012]:" | uniq

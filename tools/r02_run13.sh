#!/bin/bash
R2N2_WAITSTAT=1 timeout 300 python tools/waitstat.py --only conv1b --what fwd,dgrad 2>&1 | grep -v "^\[umma" | grep -E "==|warp  [012]:|warp  6:|warp 10:"
R2N2_FWD3_K71=1 R2N2_WAITSTAT=1 timeout 300 python tools/waitstat.py --only conv1a --what fwd 2>&1 | grep -v "^\[umma" | grep -E "==|warp  [012]:|warp  6:|warp 10:"

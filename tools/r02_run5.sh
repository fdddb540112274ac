#!/bin/bash
mkdir -p gpurun_out
R2N2_WAITSTAT=1 timeout 900 python tools/waitstat.py --out gpurun_out/r02_waitstat.txt > gpurun_out/r02_waitstat.log 2>&1; echo "rc $?" >> gpurun_out/r02_waitstat.log
tail -5 gpurun_out/r02_waitstat.log

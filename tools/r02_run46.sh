#!/bin/bash
R2N2_WAITSTAT=1 timeout 600 python tools/waitstat.py --only conv1a,conv1b --what fwd,dgrad --out gpurun_out/r02c_waitstat.txt 2>&1 | grep -v "^\[umma" | tail -60

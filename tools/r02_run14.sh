#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gputests4.log 2>&1; echo "gpu tests rc $?" >> gpurun_out/r02_gputests4.log
tail -6 gpurun_out/r02_gputests4.log

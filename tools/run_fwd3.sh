timeout 300 python -m pytest tests/test_fwd3_gpu.py -x -q > gpurun_out/pytest_f3.log 2>&1; echo exit $? >> gpurun_out/pytest_f3.log
S=conv1b
echo rot > gpurun_out/conv_f3.log; timeout 200 python tools/bench_conv.py --only $S --what fwd --iters 5 2>&1 | grep -E "TF/s" >> gpurun_out/conv_f3.log
echo norot >> gpurun_out/conv_f3.log; R2N2_FWD3_NO_ROT=1 timeout 200 python tools/bench_conv.py --only $S --what fwd --iters 5 2>&1 | grep -E "TF/s" >> gpurun_out/conv_f3.log
echo "c64 rot" >> gpurun_out/conv_f3.log; R2N2_FWD3_C64=1 timeout 200 python tools/bench_conv.py --only conv9b,conv10a --what fwd --iters 5 2>&1 | grep -E "TF/s" >> gpurun_out/conv_f3.log

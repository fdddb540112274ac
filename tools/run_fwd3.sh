timeout 300 python -m pytest tests/test_fwd3_gpu.py tests/test_rowpack_gpu.py -x -q > gpurun_out/pytest_f3.log 2>&1; echo exit $? >> gpurun_out/pytest_f3.log
R2N2_UMMA_DBG=1 timeout 200 python tools/bench_conv.py --only conv1a,conv1b,conv10b,conv11,conv11_dgrad --what fwd --iters 5 > gpurun_out/conv_f3.log 2>&1; echo exit $? >> gpurun_out/conv_f3.log
for r in 2 4; do echo "ring $r" >> gpurun_out/conv_f3.log; R2N2_FWD3_RING=$r timeout 200 python tools/bench_conv.py --only conv1a,conv1b --what fwd --iters 5 2>&1 | grep TF/s >> gpurun_out/conv_f3.log; done
for r in 4 5; do echo "ring $r" >> gpurun_out/conv_f3.log; R2N2_FWD3_RING=$r timeout 200 python tools/bench_conv.py --only conv10b,conv11 --what fwd --iters 5 2>&1 | grep TF/s >> gpurun_out/conv_f3.log; done

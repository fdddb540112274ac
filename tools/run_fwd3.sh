timeout 300 python -m pytest tests/test_fwd3_gpu.py -x -q > gpurun_out/pytest_f3.log 2>&1; echo exit $? >> gpurun_out/pytest_f3.log
R2N2_UMMA_DBG=1 timeout 200 python tools/bench_conv.py --only conv1b,conv2a,conv2b,conv9b,conv10a,conv10b,conv11,conv11_dgrad,conv9a_dgrad --what fwd --iters 5 > gpurun_out/conv_f3.log 2>&1; echo exit $? >> gpurun_out/conv_f3.log
R2N2_UMMA_NO_FWD3=1 timeout 200 python tools/bench_conv.py --only conv1b,conv2a,conv2b,conv9b,conv10a,conv10b,conv11,conv11_dgrad,conv9a_dgrad --what fwd --iters 5 > gpurun_out/conv_f3_old.log 2>&1

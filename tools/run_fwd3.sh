timeout 300 python -m pytest tests/test_fwd3_gpu.py -x -q > gpurun_out/pytest_f3.log 2>&1; echo exit $? >> gpurun_out/pytest_f3.log
S=conv1b,conv10b,conv11,conv11_dgrad
echo "heuristic" > gpurun_out/conv_f3.log; R2N2_UMMA_DBG=1 timeout 200 python tools/bench_conv.py --only $S --what fwd --iters 5 2>&1 | grep -E "TF/s|fwd3\]" | sort -u | cut -c1-330 >> gpurun_out/conv_f3.log
echo "tma epi off" >> gpurun_out/conv_f3.log; R2N2_FWD3_TMA_EPI=0 timeout 200 python tools/bench_conv.py --only $S --what fwd --iters 5 2>&1 | grep -E "TF/s" >> gpurun_out/conv_f3.log
echo "tma epi on" >> gpurun_out/conv_f3.log; R2N2_FWD3_TMA_EPI=1 timeout 200 python tools/bench_conv.py --only $S --what fwd --iters 5 2>&1 | grep -E "TF/s" >> gpurun_out/conv_f3.log

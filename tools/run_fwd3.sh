timeout 300 python -m pytest tests/test_fwd3_gpu.py tests/test_rowpack_gpu.py tests/test_kernels_gpu.py -x -q > gpurun_out/pytest_f3.log 2>&1; echo exit $? >> gpurun_out/pytest_f3.log
S=conv1a,conv1b,conv2a,conv2c
echo "default" > gpurun_out/conv_f3.log; R2N2_UMMA_DBG=1 timeout 200 python tools/bench_conv.py --only $S --what fwd --iters 5 2>&1 | grep -E "TF/s|fwd3\]|umma fwd\]" | sort -u >> gpurun_out/conv_f3.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/prof_o.json > gpurun_out/bench_o.log 2>&1

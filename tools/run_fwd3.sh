S=conv9b,conv10a
echo "fwd3 c64" > gpurun_out/conv_f3.log; R2N2_FWD3_C64=1 R2N2_UMMA_DBG=1 timeout 200 python tools/bench_conv.py --only $S --what fwd --iters 5 2>&1 | grep -E "TF/s|fwd3\]" | sort -u | cut -c1-300 >> gpurun_out/conv_f3.log
echo "persistent" >> gpurun_out/conv_f3.log; timeout 200 python tools/bench_conv.py --only $S --what fwd --iters 5 2>&1 | grep -E "TF/s" >> gpurun_out/conv_f3.log

mkdir -p gpurun_out
rm -f gpurun_out/exp_c.log
run() { echo "### $1" >> gpurun_out/exp_c.log; env $1 R2N2_UMMA_DBG=1 timeout 60 python tools/bench_conv.py --only $2 --iters 1 --what fwd 2>&1 | grep -E "umma dbg|TF/s" | tail -2 | cut -c1-260 >> gpurun_out/exp_c.log; }
run "R2N2_UMMA_SKIP=1" conv2b
run "R2N2_UMMA_SKIP=2" conv2b
run "R2N2_UMMA_SKIP=1 R2N2_UMMA_STAGES=6" conv2b
run "R2N2_UMMA_SKIP=2 R2N2_UMMA_STAGES=6" conv2b
run "R2N2_UMMA_SKIP=1" conv10b
run "R2N2_UMMA_SKIP=2" conv10b

mkdir -p gpurun_out
run() { echo "### $1" >> gpurun_out/exp_a.log; env $1 R2N2_UMMA_DBG=1 timeout 120 python tools/bench_conv.py --only conv2b --iters 1 --what fwd 2>&1 | grep -E "umma dbg|TF/s" | tail -2 >> gpurun_out/exp_a.log; }
run "X=0"
run "R2N2_UMMA_STAGES=2"
run "R2N2_UMMA_STAGES=6"
run "R2N2_UMMA_BOX=16,8,1,1"
run "R2N2_UMMA_BOX=8,16,1,1"
run "R2N2_UMMA_BOX=64,2,1,1"
run "R2N2_UMMA_SKIP=1"
run "R2N2_UMMA_SKIP=2"
run "R2N2_UMMA_SKIP=3"
run "R2N2_UMMA_SKIP=4"
run "R2N2_UMMA_SKIP=7"
run "R2N2_UMMA_BOX=16,8,1,1 R2N2_UMMA_STAGES=6"

mkdir -p gpurun_out
rm -f gpurun_out/exp_d.log
run() { echo "### $1 :: $2" >> gpurun_out/exp_d.log; env $1 R2N2_UMMA_DBG=1 timeout 60 python tools/bench_conv.py --only $2 --iters 3 --what fwd 2>&1 | grep -E "umma fwd|TF/s" | sort -u | cut -c1-200 >> gpurun_out/exp_d.log; }
S="conv1b,conv2b,conv9a,conv9b,conv10a,conv10b,conv11,conv11_dgrad,conv2c,conv9c,conv3b,conv8a"
run "R2N2_UMMA_CTAS=1" $S
run "R2N2_UMMA_CTAS=2" $S
run "R2N2_UMMA_CTAS=2 R2N2_UMMA_NSUB=1" $S
run "R2N2_UMMA_CTAS=1 R2N2_UMMA_NSUB=1" conv1b,conv2b,conv3b,conv8a
run "R2N2_UMMA_CTAS=1 R2N2_UMMA_NSUB=3" conv1b,conv2b,conv9a
run "R2N2_UMMA_CTAS=1 R2N2_UMMA_BN=128" conv3b

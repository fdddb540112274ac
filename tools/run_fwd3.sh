S=conv2a,conv2b,conv2c,conv3a,conv3b,conv4a,conv8a,conv9b,conv10a,conv9c,gru_wh_ur,conv7a
for c in "" "R2N2_UMMA_CTAS=1" "R2N2_UMMA_CTAS=2"; do echo "cfg $c" >> gpurun_out/conv_cta.log; env $c timeout 200 python tools/bench_conv.py --only $S --what fwd --iters 5 2>&1 | grep -E "TF/s" >> gpurun_out/conv_cta.log; done

#!/bin/bash
# round 2, call 4 (2 GPUs): DP equivalence on hardware, N=1 and N=2 bench on the same box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dp_equivalence_gpu.py -m gpu -q -s > gpurun_out/r02_dp_equiv_tests.log 2>&1; echo "rc $?" >> gpurun_out/r02_dp_equiv_tests.log
for c in fp32 bf16; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_equivalence.py --compute $c --out gpurun_out/r02_dp_equivalence.jsonl >> gpurun_out/r02_dp_equiv_run.log 2>&1
done
timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-infer --no-other-configs --no-parity-check > gpurun_out/r02_bench_n1.log 2> gpurun_out/r02_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline --no-infer --no-other-configs --no-parity-check > gpurun_out/r02_bench_n2.log 2> gpurun_out/r02_bench_n2.err
tail -3 gpurun_out/r02_dp_equiv_tests.log; cut -c1-400 gpurun_out/r02_bench_n1.log gpurun_out/r02_bench_n2.log

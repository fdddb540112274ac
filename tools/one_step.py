#!/usr/bin/env python
"""ONE training step of the benchmark configuration (ResidualGRUNet, B=36, T=5, bf16/tcgen05) bracketed by
cudaProfilerStart/Stop, after 3 warm-up steps: the command the per-step ncu passes wrap
(`ncu --profile-from-start off ...`).  Also writes the step's C-ABI call list (name, shape tag, algorithmic
flops / operand bytes, CUDA-event time of a separate un-profiled step) to --calls-out.

  python tools/one_step.py [--net ResidualGRUNet] [--batch 36] [--views 5] [--calls-out gpurun_out/calls.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--net', default='ResidualGRUNet')
    ap.add_argument('--batch', type=int, default=36)
    ap.add_argument('--views', type=int, default=5)
    ap.add_argument('--compute', default='bf16')
    ap.add_argument('--infer', action='store_true', help='profile one inference pass instead of a training step')
    ap.add_argument('--calls-out', default=None)
    a = ap.parse_args()
    import torch
    import r2n2_b200  # noqa: F401
    from r2n2_b200 import ops
    from r2n2_b200.lib.config import cfg
    from r2n2_b200.lib.solver import Solver
    from r2n2_b200.models import load_model
    cfg.CONST.BATCH_SIZE = a.batch
    cfg.CONST.COMPUTE = a.compute
    net = load_model(a.net)(random_seed=0)
    solver = Solver(net)
    solver.set_lr(1e-4)
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.rand(a.views, a.batch, 3, 127, 127, device='cuda', generator=g)
    occ = torch.rand(a.batch, 32, 32, 32, device='cuda', generator=g) < 0.1
    y = torch.zeros(a.batch, 32, 2, 32, 32, device='cuda')
    y[:, :, 1] = occ
    y[:, :, 0] = ~occ

    def step():
        if a.infer:
            net.engine.forward(x, None, training=False)
        else:
            solver.train_step_device(x, y)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if a.calls_out:
        ops.profiler = ops.Profiler()
        step()
        torch.cuda.synchronize()
        table, raw = ops.profiler.resolve()
        ops.profiler = None
        with open(a.calls_out, 'w') as f:
            json.dump({'sum_ms': sum(r[2] for r in raw),
                       'calls': [dict(name=n, tag=t, ms=ms, **w) for n, t, ms, w in raw]}, f, indent=0)
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    assert ops.umma_error_flag() == 0
    print('one_step OK')


if __name__ == '__main__':
    main()

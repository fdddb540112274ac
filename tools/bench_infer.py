#!/usr/bin/env python
"""Inference timings of the other BASELINE.json configurations (not the headline
bench line): CUDA-event timed Solver-level forward passes on synthetic ShapeNet-shaped inputs resident in HBM.

  config 2: GRUNet, single view, batch 36            -> ms per batch
  config 1: ResidualGRUNet, 3 views (and 1), batch 1 -> ms (demo shape)
  config 5: ResidualGRUNet, 24 views, batch 32 (= 256 / 8 GPUs) -> samples/s per GPU
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import r2n2_b200  # noqa: F401,E402
from r2n2_b200.lib.config import cfg  # noqa: E402
from r2n2_b200.models import load_model  # noqa: E402


def run(net_name, B, T, iters=10, warm=3):
    cfg.CONST.BATCH_SIZE = B
    cfg.CONST.COMPUTE = 'bf16'
    net = load_model(net_name)(random_seed=0, compute_grad=False)
    x = torch.rand(T, B, 3, 127, 127, device='cuda')
    res = {'net': net_name, 'batch': B, 'views': T}
    for name, fn in (('eager', lambda: net.engine.forward(x, None, training=False)),
                     ('cuda_graph', lambda: net.engine.forward_graph(x, None))):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res['ms_per_batch_' + name] = round(ms, 4)
        res['samples_per_s_' + name] = round(B / ms * 1e3, 1)
    return res


if __name__ == '__main__':
    out = [run('GRUNet', 36, 1), run('ResidualGRUNet', 1, 1), run('ResidualGRUNet', 1, 3), run('ResidualGRUNet', 36, 1),
           run('ResidualGRUNet', 32, 24, iters=5, warm=2), run('ResidualLSTMNet', 36, 5, iters=5, warm=2)]
    for r in out:
        print(json.dumps(r))

#!/usr/bin/env python
"""Data-parallel correctness on hardware: an N-rank step (per-rank batch b, bucketed
overlapped NCCL all-reduce of the flat gradient buffer, 1/N folded into Adam) equals a 1-rank step on the
concatenated batch N*b -- gradients and the parameters after one Adam step.

  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_equivalence.py \
      [--compute fp32|bf16] [--batch 2] [--views 2] [--out gpurun_out/dp_equivalence.json]

Every rank also builds the 1-rank reference engine locally (batch N*b), so no result has to be shipped.
fp32 path: relative L2 <= 1e-5 expected (fp32 atomics order in the weight-gradient reduction is the only
difference); bf16 path: the per-rank M tiles differ from the concatenated ones only in accumulation order.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--compute', default='fp32')
    ap.add_argument('--batch', type=int, default=2)
    ap.add_argument('--views', type=int, default=2)
    ap.add_argument('--net', default='ResidualGRUNet')
    ap.add_argument('--out', default=None)
    a = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    import r2n2_b200  # noqa: F401
    from r2n2_b200 import dist as r2dist
    from r2n2_b200.lib.config import cfg
    from r2n2_b200.lib.solver import Solver
    from r2n2_b200.models import load_model

    rank, world, local = r2dist.init_from_env()
    torch.cuda.set_device(local)
    b, T = a.batch, a.views
    cfg.CONST.COMPUTE = a.compute
    rs = np.random.RandomState(5)
    X = rs.uniform(0, 1, (T, world * b, 3, 127, 127)).astype(np.float32)
    occ = rs.rand(world * b, 32, 32, 32) < 0.1
    Y = np.zeros((world * b, 32, 2, 32, 32), np.float32)
    Y[:, :, 1] = occ
    Y[:, :, 0] = ~occ

    # N-rank data-parallel step on this rank's slice
    cfg.CONST.BATCH_SIZE = b
    net = load_model(a.net)(random_seed=0)
    solver = Solver(net)
    solver.set_lr(1e-4)
    if world > 1:
        r2dist.broadcast_params(net.engine)
        r2dist.attach(solver)
    # 1-rank step on the concatenated batch, same start
    cfg.CONST.BATCH_SIZE = world * b
    ref = load_model(a.net)(random_seed=0)
    ref.engine.store.p.copy_(net.engine.store.p)
    ref.engine.params_changed()
    rsolver = Solver(ref)
    rsolver.set_lr(1e-4)

    xs = torch.as_tensor(X[:, rank * b:(rank + 1) * b]).cuda().contiguous()
    ys = torch.as_tensor(Y[rank * b:(rank + 1) * b]).cuda().contiguous()
    out = net.forward_backward(xs, ys)
    scale = solver.grad_sync(net.engine.store.g) if solver.grad_sync is not None else 1.0
    g_dp = net.engine.store.g.double() * scale
    rout = ref.forward_backward(torch.as_tensor(X).cuda(), torch.as_tensor(Y).cuda())
    g_ref = ref.engine.store.g.double()
    loss_dp = torch.tensor([float(out['loss'])], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(loss_dp)
    loss_dp = float(loss_dp.item()) / world
    rel = lambda u, v: float((u - v).norm() / (v.norm() + 1e-300))
    res = {'world': world, 'compute': a.compute, 'per_rank_batch': b, 'views': T, 'net': a.net,
           'loss_dp_mean': loss_dp, 'loss_1rank': float(rout['loss']),
           'grad_rel_l2': rel(g_dp, g_ref), 'grad_max_abs_diff': float((g_dp - g_ref).abs().max()),
           'grad_max_abs': float(g_ref.abs().max())}
    # one optimiser step each, then compare parameters (and that every rank holds the same ones)
    solver.iteration = np.float32(1)
    solver._apply_updates(scale)
    rsolver.iteration = np.float32(1)
    rsolver._apply_updates(1.0)
    p_dp, p_ref = net.engine.store.p.double(), ref.engine.store.p.double()
    res['param_max_abs_diff_after_adam'] = float((p_dp - p_ref).abs().max())
    res['param_rel_l2_after_adam'] = rel(p_dp, p_ref)
    if world > 1:
        bits = net.engine.store.p.view(torch.int32).to(torch.int64)
        chk = torch.stack([bits.sum(), (bits * (torch.arange(bits.numel(), device=bits.device) % 8191 + 1)).sum()])
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        res['params_bit_identical_across_ranks'] = bool(all(torch.equal(c, allc[0]) for c in allc))
    if rank == 0:
        print(json.dumps(res))
        if a.out:
            with open(a.out, 'a') as f:
                f.write(json.dumps(res) + '\n')
    ok = res['grad_rel_l2'] < (1e-5 if a.compute == 'fp32' else 2e-2) and res.get('params_bit_identical_across_ranks', True)
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()

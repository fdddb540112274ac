"""world_size-2 gloo tests (CPU) of the data-parallel host logic: bucket coverage of the flat
gradient buffer, overlapped stage launches, mean-of-ranks result, 1/world scale."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import r2n2_b200  # noqa: F401
    from r2n2_b200.engine import Engine
    from r2n2_b200.dist import GradSync, broadcast_params
    eng = Engine('GRUNet', 1, device='cpu', params_only=True)
    sync = GradSync(eng)
    cov = sync.coverage()
    ok_cov = cov[0][0] == 0 and cov[-1][1] == eng.store.n and all(a[1] == b[0] for a, b in zip(cov, cov[1:]))
    g = eng.store.g
    g.copy_(torch.arange(g.numel(), dtype=torch.float32) % 97 + rank * 1000.0)
    expect = (torch.arange(g.numel(), dtype=torch.float32) % 97) * world + 1000.0 * sum(range(world))
    sync.stage_done('decoder')            # announced early (overlap), the rest at the barrier
    scale = sync(g)
    ok_sum = torch.equal(g, expect)
    eng.store.p.fill_(float(rank + 1))
    broadcast_params(eng, src=0)
    ok_bc = bool((eng.store.p == 1.0).all())
    q.put((rank, ok_cov, ok_sum, scale, ok_bc, list(sync.ranges.keys())))
    dist.destroy_process_group()


def test_gradsync_world2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_cov, ok_sum, scale, ok_bc, stages in res:
        assert ok_cov, 'buckets must tile the flat buffer exactly once'
        assert ok_sum, 'all-reduce(sum) result wrong on rank %d' % rank
        assert scale == 0.5 and ok_bc
        assert stages == ['decoder', 'rnn', 'encoder']

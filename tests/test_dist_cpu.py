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
    # split update (Solver._step): everything but the last stage is final after wait_early, the last stage after wait_late;
    # early + late ranges tile the store, and the last stage's all-reduce result is complete after wait_late
    g.copy_(torch.arange(g.numel(), dtype=torch.float32) % 89 + rank * 10.0)
    expect2 = (torch.arange(g.numel(), dtype=torch.float32) % 89) * world + 10.0 * sum(range(world))
    scale2, early, late = sync.wait_early()
    ok_early = all(torch.equal(g[lo:hi], expect2[lo:hi]) for lo, hi in early)
    tiles = sorted(early + late)
    ok_tiles = tiles[0][0] == 0 and tiles[-1][1] == g.numel() and all(a[1] == b[0] for a, b in zip(tiles, tiles[1:]))
    ok_late_is_last = sorted(late) == sorted(sync.ranges['enc_lo'])
    sync.wait_late()
    ok_all = torch.equal(g, expect2) and not sync.pending and not sync.works
    q.put((rank, ok_cov, ok_sum, scale, ok_bc, list(sync.ranges.keys()),
           (scale2, ok_early, ok_tiles, ok_late_is_last, ok_all)))
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
    for rank, ok_cov, ok_sum, scale, ok_bc, stages, split in res:
        assert split == (0.5, True, True, True, True), 'split update protocol: %r' % (split,)
        assert ok_cov, 'buckets must tile the flat buffer exactly once'
        assert ok_sum, 'all-reduce(sum) result wrong on rank %d' % rank
        assert scale == 0.5 and ok_bc
        assert stages == ['decoder', 'rnn', 'enc_hi', 'enc_mid', 'enc_lo']


def _worker_views(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ['RANK'], os.environ['WORLD_SIZE'], os.environ['LOCAL_RANK'] = str(rank), str(world), str(rank)
    import queue
    import numpy as np
    import r2n2_b200  # noqa: F401
    from r2n2_b200 import dist as r2dist
    from r2n2_b200.lib.data_process import RawViews
    r, w, _ = r2dist.init_from_env(backend='gloo')
    assert (r, w) == (rank, world)
    seed = r2dist.shared_seed(1234 + 77 * rank)          # ranks propose different seeds: rank 0's wins
    items = list(range(11))
    mine = r2dist.shard(items, rank, world)
    hq = queue.Queue()
    rs = np.random.RandomState(rank)                     # per-rank data differs
    for i in range(12):
        if i % 2:
            hq.put((RawViews(rs.randint(0, 255, (5, 2, 4, 4, 4)).astype(np.uint8), np.zeros((5, 2, 6), np.int32)),
                    np.zeros((2, 4, 4, 4), np.uint8)))
        else:
            hq.put((rs.rand(5, 2, 3, 4, 4).astype(np.float32), np.zeros((2, 4, 2, 4, 4), np.float32)))
    sq = r2dist.SharedViewCountQueue(hq, seed, 5)
    shapes = []
    for i in range(12):
        x, y = sq.get()
        shapes.append(int(x.shape[0]))
        assert (hasattr(x, 'rgba')) == bool(i % 2) and x.shape[1] == 2
    mean = r2dist.all_reduce_mean_scalar(float(rank + 1))
    q.put((rank, seed, mine, shapes, list(sq.counts), mean))
    dist.destroy_process_group()


def test_shared_view_schedule_and_dataset_shards_world2_gloo():
    """lib/train_net.train_net() under a launcher: disjoint covering dataset shards, ONE view-count schedule on
    all ranks (reference draw: lib/data_process.py:127-130), rank 0's seed everywhere."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_views, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, seed0, mine0, shapes0, counts0, mean0), (r1, seed1, mine1, shapes1, counts1, mean1) = res
    assert seed0 == seed1 == 1234
    assert sorted(mine0 + mine1) == list(range(11)) and not set(mine0) & set(mine1)
    assert shapes0 == shapes1 == counts0 == counts1
    assert all(1 <= t <= 5 for t in shapes0) and len(set(shapes0)) > 1
    assert mean0 == mean1 == 1.5


def test_numa_binding_is_a_noop_without_cuda():
    """dist.bind_to_gpu_numa only ever narrows the affinity to the GPU's local cores; without CUDA (or when the
    topology cannot be read) it returns None and leaves the process affinity alone."""
    sys.path.insert(0, ROOT)
    import r2n2_b200  # noqa: F401
    from r2n2_b200 import dist as r2dist
    before = os.sched_getaffinity(0)
    assert r2dist.bind_to_gpu_numa(0) is None or torch.cuda.is_available()
    if not torch.cuda.is_available():
        assert os.sched_getaffinity(0) == before

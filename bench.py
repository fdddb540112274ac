#!/usr/bin/env python
"""Benchmark of the 3D-R2N2 hot path on B200 (contract: see the task description / DESIGN.md §5).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--compute bf16|fp32] [--batch 36] [--views 5] [--net ResidualGRUNet]

Workload = BASELINE.json config[2], the configuration the headline metric is quoted on:
ResidualGRUNet training (SoftmaxWithLoss3D + Adam), per-GPU batch 36, 5 views of 127x127 RGB ->
32^3x2, synthetic ShapeNet-shaped data, random-init weights.  One "step" = forward + backward +
(N>1: NCCL gradient all-reduce) + fused Adam over one batch.

Printed JSON (rank 0, one line):
  value   : samples/s, whole job, inputs resident in HBM (CUDA-event timed, max over ranks)
  e2e     : same metric through Solver.train_loss with HOST (pinned) numpy inputs: H2D of x,y and
            D2H of the loss inside the timed region
  e2e_raw_pipeline: same, fed by lib/data_process.DeviceBatchQueue from host uint8 RGBA renders (what
            lib/train_net.train_net() runs): H2D of 14.7 MB instead of 44.3 MB, preprocessing kernels included
  infer   : 1-view inference latencies (the second half of BASELINE.json's metric), CUDA-graph replay
  roofline: dominant kernel (implicit-GEMM conv fwd/dgrad), algorithmic (useful) FLOPs / CUDA-event
            time of its launches, against MEASURED_PEAKS.json
  cpu_baseline: the oracle's torch-CPU restatement of the reference graph timed on the host cores
            on a bounded sample (kind "port": Theano itself cannot run)
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'ResidualGRUNet train samples/sec (5 views, 32^3)'
UNIT = 'samples/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--compute', default=os.environ.get('R2N2_BENCH_COMPUTE', 'auto'))
    ap.add_argument('--net', default='ResidualGRUNet')
    ap.add_argument('--batch', type=int, default=36)
    ap.add_argument('--views', type=int, default=5)
    ap.add_argument('--cpu-sample-batch', type=int, default=2)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-infer', action='store_true')
    ap.add_argument('--no-raw-pipeline', action='store_true')
    ap.add_argument('--no-parity-check', action='store_true')
    ap.add_argument('--no-other-configs', action='store_true', help='skip the config 4 / config 5 legs')
    ap.add_argument('--profile-out', default=None, help='write the per-kernel event table here')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=d.get('bf16_tflops_sustained', d.get('bf16_tflops')), hbm=d.get('hbm_gbs'),
                    src='measured (MEASURED_PEAKS.json, bf16 sustained)')
    return dict(tflops=1400.0, hbm=6650.0, src='fallback (B200_PROFILING.md)')


def ncu_traffic(n_launches):
    """DRAM bytes per launch of the dominant kernel family: sum of dram__bytes_read.sum + dram__bytes_write.sum
    over ALL forward / data-gradient conv launches of one training step (the same launches `roofline.frac`
    covers), divided by their count.  From the committed ncu pass (profiles/ncu_traffic.json, written by
    tools/ncu_traffic.py from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` over one step)."""
    p = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if not os.path.exists(p):
        return None, None
    try:
        d = json.load(open(p)).get('conv_fwd_step')
        if not d:
            return None, None
        return d['dram_bytes_total'] / max(1, d['launches']), d
    except Exception:
        return None, None


def config_dict(args, extra=None):
    c = {'workload': '%s training step (fwd+bwd+Adam), per-GPU batch %d, %d views 127x127 RGB -> 32^3x2'
                     % (args.net, args.batch, args.views),
         'net': args.net, 'per_gpu_batch': args.batch, 'views': args.views,
         'global_batch': args.batch * args.gpus, 'parallelism': 'dp%d' % args.gpus,
         'l2': 'working set per step (activations, GBs) >> 126 MB L2; no explicit flush'}
    if extra:
        c.update(extra)
    return c


def synthetic_batch(np, T, B, seed_x=1234, seed_y=4321, n_vox=32, img=127):
    """ShapeNet-shaped synthetic inputs (lib/data_process.py:133-154 shapes/dtypes/ranges):
    x (T,B,3,127,127) uniform [0,1); y (B,32,2,32,32) one-hot occupancy, ~10% occupied."""
    x = np.random.RandomState(seed_x).uniform(0, 1, (T, B, 3, img, img)).astype(np.float32)
    occ = np.random.RandomState(seed_y).rand(B, n_vox, n_vox, n_vox) < 0.1
    y = np.zeros((B, n_vox, 2, n_vox, n_vox), np.float32)
    y[:, :, 1] = occ
    y[:, :, 0] = ~occ
    return x, y


# ------------------------------------------------------------------------------- CPU reference
def cpu_reference_step(net_name, B, T, n_steps, warmup, min_seconds=0.0):
    """Times training steps of the oracle's torch-CPU restatement of the reference graph
    (oracle/ref_torch.py: per-view encoder inside the loop, dense convs on zero-stuffed tensors).
    min_seconds > 0: keep stepping past n_steps until that much timed CPU work has been done."""
    import torch
    from oracle import ref_torch as rt
    torch.set_num_threads(os.cpu_count() or 1)
    params = rt.init_params(net_name, seed=0)
    x, y = rt.synthetic_batch(T, B)
    net = rt.RefNet(net_name, params)
    adam = rt.RefAdam(net, lr=1e-4)
    times = []
    i = 0
    while i < warmup + n_steps or (sum(times) < min_seconds and len(times) < 64):
        t0 = time.perf_counter()
        adam.step(x, y)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        i += 1
    return times, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    B = args.cpu_sample_batch
    steps = max(1, min(args.steps, 240))          # K steps as asked (0.5 s each at batch 2); capped at ~2 minutes
    warm = max(0, min(args.warmup, 10))
    times, cores = cpu_reference_step(args.net, B, args.views, steps, warm)
    ms = 1e3 * sum(times) / len(times)
    val = B / (ms / 1e3)
    sample = ('oracle torch-CPU fp32 restatement of the reference graph, %d training steps at batch %d '
              '(same net/views, batch reduced from %d to bound the CPU time)' % (steps, B, args.batch))
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': steps, 'warmup': warm, 'ms_per_step': ms, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': config_dict(args, {'cpu_sample_batch': B}),
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.p = None
        self.idx = gpu_index
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '50'],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            return None
        sm, mx, reasons = [], [], set()
        for ln in out.strip().splitlines():
            f = [t.strip() for t in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return None
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': max(mx), 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------------- ours
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import r2n2_b200  # noqa: F401
    from r2n2_b200 import ops, _lib
    from r2n2_b200.lib.config import cfg
    from r2n2_b200.lib.solver import Solver
    from r2n2_b200.models import load_model
    from r2n2_b200 import dist as r2dist

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    if world > 1:
        r2dist.init_from_env('nccl')        # (also caps the SMs NCCL may take: NCCL_MAX_CTAS, see dist.py)
    _lib.load()

    compute = args.compute
    if compute == 'auto':
        compute = os.environ.get('R2N2_COMPUTE', 'bf16')
    B, T = args.batch, args.views
    cfg.CONST.BATCH_SIZE = B
    cfg.CONST.COMPUTE = compute
    net = load_model(args.net)(random_seed=0)
    solver = Solver(net)
    solver.set_lr(cfg.TRAIN.DEFAULT_LEARNING_RATE)
    if world > 1:
        r2dist.broadcast_params(net.engine)

    # synthetic ShapeNet-shaped batch (different per rank), pinned host copies for the e2e leg
    x_np, y_np = synthetic_batch(np, T, B, seed_x=1234 + rank, seed_y=4321 + rank)
    x_pin = torch.as_tensor(x_np).pin_memory()
    y_pin = torch.as_tensor(y_np).pin_memory()
    x_dev, y_dev = x_pin.cuda(), y_pin.cuda()

    # ---- parity check on the timed batch (outside every timed region): the benchmarked arithmetic against the
    # exact-fp32 FFMA path of the same library on the same weights and inputs (that path is pinned to the CPU
    # oracle at 3e-6 by tests/test_nets_gpu.py).  One fp32 forward + backward (~0.4 s) per rank.
    parity = None
    if compute != 'fp32' and not args.no_parity_check:
        cfg.CONST.COMPUTE = 'fp32'
        net32 = load_model(args.net)(random_seed=0)
        cfg.CONST.COMPUTE = compute
        net32.engine.set_params(net.engine.get_params())
        o32 = net32.engine.forward(x_dev, y_dev, training=True, want_pred=True)
        net32.engine.backward()
        p32, l32 = o32['pred'].clone(), float(o32['loss'])
        o16 = net.engine.forward(x_dev, y_dev, training=True, want_pred=True)
        net.engine.backward()
        g16, g32 = net.engine.get_grads(), net32.engine.get_grads()
        num = sum(float(((a.astype(np.float64) - b) ** 2).sum()) for a, b in zip(g16, g32))
        den = sum(float((b.astype(np.float64) ** 2).sum()) for b in g32)
        thr = 0.4
        occ = y_dev[:, :, 1] > 0.5

        def iou(p):
            q = p[:, :, 1] > thr
            return float((q & occ).sum()) / max(1.0, float((q | occ).sum()))
        parity = {'loss_bf16': float(o16['loss']), 'loss_fp32': l32,
                  'max_abs_dp': float((o16['pred'] - p32).abs().max()),
                  'mean_abs_dp': float((o16['pred'] - p32).abs().mean()),
                  'grad_rel_l2': (num / den) ** 0.5, 'iou04_bf16': iou(o16['pred']), 'iou04_fp32': iou(p32),
                  'against': 'fp32 FFMA path of the same library (oracle-pinned), same weights, timed batch of rank 0'}
        del net32, o32, p32, g16, g32
        torch.cuda.empty_cache()
    if world > 1:
        r2dist.attach(solver)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        sampler = ClockSampler(local) if rank == 0 else None
        l0 = ops.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = ops.launch_count - l0
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([ms], device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), launches, clocks

    # ---- value: inputs resident in HBM ------------------------------------------------------
    losses = []

    def step_dev():
        losses.append(solver.train_step_device(x_dev, y_dev))
    ms_total, launches, clocks = timed(step_dev, args.steps, args.warmup)
    ms_step = ms_total / args.steps
    value = B * world / (ms_step / 1e3)
    final_loss = float(losses[-1])

    # ---- e2e: host buffers through the public API (Solver.train_loss) -----------------------
    def step_e2e():
        solver.train_loss(x_pin, y_pin)        # H2D x,y (pinned) ... D2H loss (float())
    e2e_steps = max(2, min(args.steps, 20))
    ms_e2e_total, _, _ = timed(step_e2e, e2e_steps, 1)
    e2e_value = B * world / (ms_e2e_total / e2e_steps / 1e3)

    # ---- e2e through the input pipeline: raw uint8 renders -> DeviceBatchQueue -> train_loss ----
    # (what lib/train_net.train_net() runs: the H2D copy carries 13.5 MB of RGBA bytes instead of 44 MB of
    # floats, the crop / flip / background / scale and the label expansion are kernels on a side stream one
    # batch ahead; the workers' PNG decode is not part of it)
    e2e_raw = None
    if not args.no_raw_pipeline:
        import queue as _q
        from r2n2_b200.lib import data_process as dp
        rr = np.random.RandomState(77 + rank)
        raw_steps = max(e2e_steps, min(args.steps, 20))
        n_raw = raw_steps + 3
        raw = []
        for _ in range(2):                                    # two distinct batches, reused in turn
            rgba = rr.randint(0, 256, (T, B, 137, 137, 4)).astype(np.uint8)
            params = np.zeros((T, B, 6), np.int32)
            params[..., 0:2] = rr.randint(0, 10, (T, B, 2))
            params[..., 2] = rr.randint(0, 2, (T, B))
            params[..., 3:6] = rr.randint(225, 256, (T, B, 3))
            occ = (rr.rand(B, 32, 32, 32) < 0.1).astype(np.uint8)
            raw.append((dp.RawViews(rgba, params), occ))
        hq = _q.Queue()
        for i in range(n_raw):
            hq.put(raw[i % 2])
        dq = dp.DeviceBatchQueue(hq, depth=2)

        def step_raw():
            xb, yb = dq.get(timeout=120)
            solver.train_loss(xb, yb)
        ms_raw_total, _, _ = timed(step_raw, raw_steps, 3)
        dq.close()
        e2e_raw = {'value': B * world / (ms_raw_total / raw_steps / 1e3), 'unit': UNIT,
                   'h2d_bytes_per_step': int(raw[0][0].rgba.nbytes + raw[0][0].params.nbytes + raw[0][1].nbytes),
                   'd2h_bytes_per_step': 4, 'steps': raw_steps,
                   'note': 'Solver.train_loss fed by lib/data_process.DeviceBatchQueue from host uint8 RGBA renders'}

    # ---- roofline pass: CUDA events around every C-ABI launch of one more step (rank 0) ------
    roofline, table_out = None, None
    pk = peaks()
    # every rank runs the step (it contains the gradient all-reduce); only rank 0 records events
    # PROFILE_REPS steps, each launch's time = the median over the steps (one profiled step alone moves the fraction by
    # +-3 % from run to run on a power-capped box)
    PROFILE_REPS = 5
    raws = []
    for _ in range(PROFILE_REPS):
        prof = ops.Profiler() if rank == 0 else None
        ops.profiler = prof
        step_dev()
        torch.cuda.synchronize()
        ops.profiler = None
        if rank == 0:
            raws.append(prof.resolve()[1])
    if rank == 0:
        same = all(len(r_) == len(raws[0]) and all(a[0] == b[0] for a, b in zip(r_, raws[0])) for r_ in raws)
        if not same:
            raws = raws[:1]
        raw, table = [], {}
        for i, (name, tag, _ms, work) in enumerate(raws[0]):
            ms = float(np.median([r_[i][2] for r_ in raws]))
            raw.append((name, tag, ms, work))
            d = table.setdefault(name, dict(ms=0.0, calls=0, flops=0.0, bytes=0.0))
            d['ms'] += ms
            d['calls'] += 1
            d['flops'] += work.get('flops', 0.0)
            d['bytes'] += work.get('bytes', 0.0)
        tot_ms = sum(d['ms'] for d in table.values())
        # forward / data-gradient launches of the implicit-GEMM kernels (the _colsum entry point is the same
        # kernels plus the producing layer's bias gradient, fused or as a second pass: counted in, conservatively)
        conv = dict(ms=0.0, flops=0.0, calls=0, bytes=0.0)
        for name in ('r2n2_conv_fwd', 'r2n2_conv_fwd_colsum', 'r2n2_conv_fwd_gru'):
            for k_ in conv:
                conv[k_] += table.get(name, {}).get(k_, 0)
        if conv['ms'] > 0:
            ach = conv['flops'] / (conv['ms'] * 1e-3) / 1e12
            traffic, tdetail = ncu_traffic(conv['calls'])
            roofline = {'bound': 'tensor', 'kernel': 'r2n2_conv_fwd / _colsum / _gru (implicit-GEMM conv/FC fwd + dgrad)',
                        'achieved': ach, 'peak': pk['tflops'], 'unit': 'TFLOP/s', 'frac': ach / pk['tflops'],
                        'traffic': traffic,
                        'traffic_note': None if tdetail is None else
                        ('ncu dram bytes summed over the %d conv fwd/dgrad kernel launches of one step / %d; '
                         'algorithmic operand bytes of the same launches: %.0f per launch'
                         % (tdetail['launches'], tdetail['launches'],
                            conv['bytes'] / max(1, conv['calls']))),
                        'peak_source': pk['src'],
                        'launches': conv['calls'], 'avg_launch_ms': conv['ms'] / max(1, conv['calls']),
                        'share_of_step': conv['ms'] / tot_ms if tot_ms else None,
                        'flops_counted': 'algorithmic (useful, zero-skipped unpool), summed over the launches',
                        'timing': 'CUDA events around every launch, per-launch median over %d profiled steps' % len(raws)}
        table_out = {k: dict(v, share=v['ms'] / tot_ms if tot_ms else 0.0,
                             tflops=(v['flops'] / (v['ms'] * 1e-3) / 1e12) if v['ms'] > 0 else 0.0,
                             gbs=(v['bytes'] / (v['ms'] * 1e-3) / 1e9) if v['ms'] > 0 else 0.0)
                     for k, v in sorted(table.items(), key=lambda kv: -kv[1]['ms'])}
        if args.profile_out:
            with open(args.profile_out, 'w') as f:
                json.dump({'per_kernel': table_out, 'sum_ms': tot_ms,
                           'launches': [(n, t, ms, w) for n, t, ms, w in raw]}, f, indent=1)

    # ---- data-parallel consistency: after all those steps every rank must hold bit-identical parameters ----
    dp_check = None
    if world > 1:
        st = net.engine.store
        # where does the step time beyond the 1-GPU figure go?  Each rank's own forward+backward+Adam with the gradient
        # synchronisation switched OFF (parameters restored afterwards): the slowest of N GPUs sets the synchronous step
        # (GPUs of one box differ by a few % under the power cap); what is left over is communication.
        p_keep, m_keep, v_keep = st.p.clone(), None if st.m is None else st.m.clone(), None if st.v is None else st.v.clone()
        it_keep = solver.iteration
        sync_keep, prog_keep, hook_keep = solver.grad_sync, net.engine.progress, ops.pre_launch
        solver.grad_sync, net.engine.progress, ops.pre_launch = None, None, None
        for _ in range(3):
            solver.train_step_device(x_dev, y_dev)
        torch.cuda.synchronize()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        for _ in range(10):
            solver.train_step_device(x_dev, y_dev)
        eb.record()
        torch.cuda.synchronize()
        own = torch.tensor([ea.elapsed_time(eb) / 10.0], device='cuda')
        solver.grad_sync, net.engine.progress, ops.pre_launch = sync_keep, prog_keep, hook_keep
        st.p.copy_(p_keep)
        if m_keep is not None:
            st.m.copy_(m_keep), st.v.copy_(v_keep)
        solver.iteration = it_keep
        net.engine.params_changed()
        owns = [torch.zeros_like(own) for _ in range(world)]
        dist.all_gather(owns, own)
        own_ms = [round(float(o.item()), 3) for o in owns]
        bits = st.p.view(torch.int32).to(torch.int64)
        chk = torch.stack([bits.sum(), (bits * (torch.arange(bits.numel(), device=bits.device) % 8191 + 1)).sum()])
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        dp_check = {'params_bit_identical_across_ranks': bool(all(torch.equal(c, allc[0]) for c in allc)),
                    'checksum': [int(v) for v in allc[0].tolist()], 'steps_taken': int(solver.iteration),
                    'per_rank_ms_without_sync': own_ms, 'slowest_rank_ms': max(own_ms),
                    'sync_overhead_ms': round(ms_step - max(own_ms), 3),
                    'note': 'ms_per_step = slowest rank + sync_overhead; the N=1 figure is one GPU of the box'}

    # ---- CPU baseline (rank 0, N=1 only, bounded sample) --------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Bc = args.cpu_sample_batch
        times, cores = cpu_reference_step(args.net, Bc, T, 3, 1, min_seconds=10.0)
        v = Bc / (sum(times) / len(times))
        cpu = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
               'sample': 'oracle torch-CPU fp32 restatement of the reference graph: 1 warm-up + %d timed training '
                         'steps (%.1f s of CPU work) at batch %d, %d views (batch reduced from %d)'
                         % (len(times), sum(times), Bc, T, B)}

    # ---- 1-view inference latency (second half of BASELINE.json's metric; rank 0, N=1 only) ---------
    infer = None
    if rank == 0 and world == 1 and not args.no_infer:
        infer = {}
        for name, net_name, Bi in (('GRUNet_b36_1view_ms', 'GRUNet', 36), ('ResidualGRUNet_b1_1view_ms', 'ResidualGRUNet', 1),
                                   ('ResidualGRUNet_b36_1view_ms', 'ResidualGRUNet', 36)):
            cfg.CONST.BATCH_SIZE = Bi
            neti = load_model(net_name)(random_seed=0, compute_grad=False)
            xi = torch.rand(1, Bi, 3, 127, 127, device='cuda')
            for _ in range(3):
                neti.engine.forward_graph(xi, None)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                neti.engine.forward_graph(xi, None)
            b.record()
            torch.cuda.synchronize()
            infer[name] = round(a.elapsed_time(b) / 20, 4)
            del neti
        cfg.CONST.BATCH_SIZE = B
        infer['note'] = 'bf16/tcgen05 forward replayed as one CUDA graph, input resident in HBM, CUDA events'

    # ---- BASELINE.json configs 4 and 5 at every N (weak scaling, same per-GPU shapes) ---------------------
    other = None
    if not args.no_other_configs:
        other = {}
        del solver, net
        torch.cuda.empty_cache()
        # config 4: 3D-conv-LSTM variant of the residual net, 5-view training, batch 36 per GPU
        cfg.CONST.BATCH_SIZE = B
        net4 = load_model('ResidualLSTMNet')(random_seed=0)
        s4 = Solver(net4)
        s4.set_lr(cfg.TRAIN.DEFAULT_LEARNING_RATE)
        if world > 1:
            r2dist.broadcast_params(net4.engine)
            r2dist.attach(s4)
        k4 = max(3, min(args.steps, 10))
        ms4, _, _ = timed(lambda: s4.train_step_device(x_dev, y_dev), k4, 3)
        other['lstm_train'] = {'metric': 'ResidualLSTMNet train samples/sec (5 views, 32^3)',
                               'value': B * world / (ms4 / k4 / 1e3), 'unit': UNIT, 'ms_per_step': ms4 / k4,
                               'steps': k4, 'warmup': 3, 'per_gpu_batch': B, 'views': T, 'n_gpus': world,
                               'config': 'BASELINE.json configs[3]'}
        del s4, net4
        torch.cuda.empty_cache()
        # config 5: ResidualGRUNet 24-view sequence inference, batch 256 on 8 GPUs = 32 per GPU, batch-sharded, no
        # collective; one CUDA-graph replay per step, input resident in HBM
        B5, T5 = 32, 24
        cfg.CONST.BATCH_SIZE = B5
        net5 = load_model('ResidualGRUNet')(random_seed=0, compute_grad=False)
        x5 = torch.rand(T5, B5, 3, 127, 127, device='cuda')
        k5 = max(3, min(args.steps, 10))
        ms5, _, _ = timed(lambda: net5.engine.forward_graph(x5, None), k5, 3)
        other['infer_24view'] = {'metric': 'ResidualGRUNet 24-view inference samples/sec (batch 32 per GPU)',
                                 'value': B5 * world / (ms5 / k5 / 1e3), 'unit': UNIT, 'ms_per_batch': ms5 / k5,
                                 'steps': k5, 'warmup': 3, 'per_gpu_batch': B5, 'views': T5, 'n_gpus': world,
                                 'config': 'BASELINE.json configs[4]'}
        del net5, x5
        cfg.CONST.BATCH_SIZE = B

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None,
                'dtype': {'bf16': 'bf16', 'bf16x3': 'bf16x3 (hi/lo operand split, fp32 accumulate and storage)'}.get(compute, 'f32'),
                'data': 'synthetic',
                'config': config_dict(args, {'compute': compute,
                                             'conv_algo': 'tcgen05/TMA' if compute.startswith('bf16') else 'fp32 FFMA',
                                             'final_loss': final_loss}),
                'e2e': {'value': e2e_value, 'unit': UNIT,
                        'h2d_bytes_per_step': int(x_pin.numel() * 4 + y_pin.numel() * 4),
                        'd2h_bytes_per_step': 4, 'steps': e2e_steps},
                'e2e_raw_pipeline': e2e_raw,
                'gpu_launches': launches, 'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu,
                'infer': infer, 'parity_check': parity, 'other_configs': other, 'dp_check': dp_check,
                'per_kernel': {k: {'ms': round(v['ms'], 4), 'calls': v['calls'], 'share': round(v['share'], 4),
                                   'tflops': round(v['tflops'], 2), 'gbs': round(v['gbs'], 1)}
                               for k, v in (table_out or {}).items()}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()

"""Data-parallel training support: one process per GPU, equal per-rank batch,
gradient all-reduce(sum) x 1/world over the flat gradient buffer == gradient of the global-batch
mean loss (lib/layers.py:600).  The reference itself is single-device (main.py:77).

The flat buffer is reduced in a few large buckets, launched as soon as the backward pass has
finished writing them (decoder -> recurrent unit + fc7 -> encoder + biases), so that the NCCL
traffic over NVLink overlaps the remaining backward kernels; the 1/world scale is folded into
the fused optimiser kernel (r2n2_adam's grad_scale).  No other collective exists on the path;
inference shards by batch with no communication at all.
"""
import torch
import torch.distributed as dist


def bucket_ranges(store, stage_names):
    """Split [0, store.n) into contiguous element ranges, one per backward stage.
    stage_names: {stage: [segment names]} for the non-bias segments; biases (the tail of the
    buffer, from store.n_decay on) go with the last stage.  Returns {stage: [(lo, hi), ...]}."""
    out = {}
    for stage, names in stage_names.items():
        spans = sorted((store.offsets[n][0], store.offsets[n][0] + ((store.offsets[n][1] + 63) // 64) * 64)
                       for n in names)
        merged = []
        for lo, hi in spans:
            if merged and merged[-1][1] == lo:
                merged[-1] = (merged[-1][0], hi)
            else:
                merged.append((lo, hi))
        out[stage] = merged
    return out


class GradSync:
    """Bucketed, overlapped all-reduce of engine.store.g.

    engine.backward(progress=sync.stage_done) calls stage_done(stage) when a stage's gradients
    are final; __call__ (used as Solver.grad_sync) waits for all buckets and returns the scale
    the optimiser must apply (1/world)."""

    def __init__(self, engine, group=None):
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        st = engine.store
        stages = engine.stage_segments()          # ordered dict-like: decoder, rnn, encoder
        self.ranges = bucket_ranges(st, stages)
        last = list(stages.keys())[-1]
        if st.n > st.n_decay:
            self.ranges[last] = self.ranges[last] + [(st.n_decay, st.n)]
        self.pending = []
        self.launched = set()
        self.works = {}                # stage -> its all-reduce handles
        self.last_stage = last
        # OPTIONAL SM partition while an all-reduce is in flight (csrc: r2n2_set_sm_limit): the compute grids leave
        # `reserve` TPCs (2 SMs each: the CTA-pair kernels need whole TPCs) to the NCCL kernels, which init_from_env caps
        # at NCCL_MAX_CTAS = reserve channels.  Measured on 8 x B200 (profiles/r02_scaling_experiments.txt): no gain --
        # 11.52 ms/step without, 11.56 / 11.59 / 11.87 ms with 4 / 8 / 16 reserved -- so it is OFF by default
        # (R2N2_COMM_SMS=0); the step-time gap to N=1 is the slowest GPU of the box, see bench.py dp_check.
        import os
        self.reserve = int(os.environ.get('R2N2_COMM_SMS', '0')) if st.p.is_cuda else 0
        self.n_sm = torch.cuda.get_device_properties(st.p.device).multi_processor_count if st.p.is_cuda else 0
        self._limited = False

    def _set_limit(self, on):
        if self.reserve <= 0 or on == self._limited:
            return
        from . import _lib as L
        L.call('r2n2_set_sm_limit', self.n_sm - 2 * self.reserve if on else 0)
        self._limited = on

    def poll(self):
        """ops.pre_launch hook: give the SMs back as soon as every launched all-reduce has completed."""
        if self._limited and all(w.is_completed() for w in self.pending):
            self._set_limit(False)

    def coverage(self):
        """all (lo, hi) ranges, sorted: must tile [0, n) exactly once (checked in tests)."""
        return sorted(r for rs in self.ranges.values() for r in rs)

    def stage_done(self, stage):
        if self.world == 1 or stage in self.launched:
            return
        self.launched.add(stage)
        g = self.engine.store.g
        ws = [dist.all_reduce(g[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
              for lo, hi in self.ranges[stage]]
        self.works[stage] = ws
        self.pending.extend(ws)
        self._set_limit(True)

    def __call__(self, g):
        if self.world > 1:
            for stage in self.ranges:              # anything the backward did not announce
                self.stage_done(stage)
            for w in self.pending:
                w.wait()
        self.pending, self.launched, self.works = [], set(), {}
        self._set_limit(False)
        return 1.0 / self.world

    # The last stage's buckets (conv1a / conv1b and all biases, 0.4 MB) are launched after the final backward kernel: nothing
    # of the backward pass is left to hide them.  OPT-IN (R2N2_SPLIT_UPDATE=1): Solver runs the optimiser in two passes --
    # everything else right away (wait_early), the last stage's ranges once their all-reduce has landed (wait_late).
    # Measured slower than the single pass (see Solver._step), kept as a tested experiment.
    def wait_early(self):
        """-> (scale, early_ranges, late_ranges); the last stage's all-reduce may still be in flight."""
        n = self.engine.store.p.numel()
        if self.world == 1:
            return 1.0, [(0, n)], []
        for stage in self.ranges:
            self.stage_done(stage)
        late = sorted(self.ranges[self.last_stage])
        for stage, ws in self.works.items():
            if stage != self.last_stage:
                for w in ws:
                    w.wait()
        early, pos = [], 0
        for lo, hi in late:
            if lo > pos:
                early.append((pos, lo))
            pos = hi
        if pos < n:
            early.append((pos, n))
        return 1.0 / self.world, early, late

    def wait_late(self):
        for w in self.works.get(self.last_stage, []):
            w.wait()
        self.pending, self.launched, self.works = [], set(), {}
        self._set_limit(False)


def attach(solver, group=None):
    """Make `solver` data-parallel over the default (or given) process group."""
    sync = GradSync(solver.net.engine, group)
    solver.grad_sync = sync
    solver.net.engine.progress = sync.stage_done
    if sync.world > 1 and sync.reserve > 0:
        from . import ops
        ops.pre_launch = sync.poll
    solver.rank = dist.get_rank(group) if dist.is_initialized() else 0
    solver.world = sync.world
    solver.dist_group = group
    return sync


def init_from_env(backend=None):
    """Join the process group a launcher (torchrun: RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*) describes.
    Returns (rank, world, local_rank); (0, 1, 0) and no group when the process was started alone."""
    import os
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world <= 1:
        return 0, 1, 0
    rank, local = int(os.environ['RANK']), int(os.environ.get('LOCAL_RANK', '0'))
    bind_to_gpu_numa(local)
    # SM partition between the step's persistent kernels and the overlapped all-reduce (GradSync._set_limit):
    # NCCL gets at most NCCL_MAX_CTAS channels (= SMs); must be set before the communicator exists; explicit settings win.
    reserve = int(os.environ.get('R2N2_COMM_SMS', '0'))
    if reserve > 0 and torch.cuda.is_available():
        os.environ.setdefault('NCCL_MAX_CTAS', str(reserve))
    if not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
            dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def bind_to_gpu_numa(local_rank):
    """Pin this process (and the threads it starts later: data staging, workers) to the CPU cores of the NUMA node
    the GPU hangs off (sysfs `local_cpulist` of the GPU's PCI function).  With one process per GPU on a two-socket
    host the pinned staging buffers are then first-touched on the GPU's own socket and the per-step host-to-device
    copies do not cross the inter-socket link.  Returns the core set, or None when the topology cannot be read
    (single-socket host, container without sysfs, R2N2_NO_NUMA_BIND=1): nothing is changed then."""
    import os
    if os.environ.get('R2N2_NO_NUMA_BIND') or not torch.cuda.is_available() or not hasattr(os, 'sched_setaffinity'):
        return None
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = '%04x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open('/sys/bus/pci/devices/%s/local_cpulist' % bdf) as fh:
            text = fh.read().strip()
        cpus = set()
        for part in text.split(','):
            if not part:
                continue
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def shard(items, rank, world):
    """Rank `rank`'s share of a dataset list: every world-th item (disjoint, covering, sizes differ by <= 1)."""
    return list(items)[rank::world]


def shared_seed(seed, group=None):
    """Rank 0's `seed` on every rank (one small broadcast at start-up)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return int(seed)
    box = [int(seed)]
    dist.broadcast_object_list(box, src=0, group=group)
    return int(box[0])


class SharedViewCountQueue(object):
    """Queue wrapper that gives every rank the SAME number of views per step.

    The reference draws the view count of a batch inside the data worker (lib/data_process.py:127-130:
    `np.random.randint(n_views) + 1`, then `np.random.choice(NUM_RENDERING, curr_n_views)` per sample).  With
    one process per GPU the ranks' workers would draw different counts, and the recurrence length must agree
    across ranks for the bucketed gradient all-reduce to line up step by step.  Here the workers always decode
    N_VIEWS renders (cfg.TRAIN.RANDOM_NUM_VIEWS off in the workers) and the trainer keeps the first T_k of them,
    T_k ~ U{1..N_VIEWS} drawn from a RandomState seeded with one broadcast seed: the first T of N i.i.d.
    rendering ids are T i.i.d. rendering ids, so the distribution of a rank's batch is the reference's."""

    def __init__(self, data_queue, seed, n_views, random_num_views=True):
        import numpy as np
        self.data_queue = data_queue
        self.n_views = int(n_views)
        self.random_num_views = random_num_views
        self._rs = np.random.RandomState(int(seed) % (2 ** 32))
        self.counts = []                    # drawn schedule (tests / logging)

    def next_count(self):
        t = int(self._rs.randint(self.n_views)) + 1 if self.random_num_views else self.n_views
        self.counts.append(t)
        return t

    def get(self, *args, **kwargs):
        batch_img, batch_voxel = self.data_queue.get(*args, **kwargs)
        t = self.next_count()
        have = batch_img.shape[0]
        if have < t:
            raise ValueError('worker batch holds %d views, the shared schedule asks for %d '
                             '(workers must run with TRAIN.RANDOM_NUM_VIEWS off)' % (have, t))
        if hasattr(batch_img, 'rgba'):      # lib/data_process.RawViews
            batch_img = type(batch_img)(batch_img.rgba[:t], batch_img.params[:t])
        else:
            batch_img = batch_img[:t]
        return batch_img, batch_voxel

    def empty(self):
        return self.data_queue.empty()


def all_reduce_mean_scalar(value, group=None):
    """Mean over ranks of a host scalar (validation loss)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    dev = 'cuda' if dist.get_backend(group) == 'nccl' else 'cpu'
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item()) / dist.get_world_size(group)


def broadcast_params(engine, src=0, group=None):
    """Start all ranks from rank `src`'s parameters."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(engine.store.p, src=src, group=group)
        engine.params_changed()

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

    def coverage(self):
        """all (lo, hi) ranges, sorted: must tile [0, n) exactly once (checked in tests)."""
        return sorted(r for rs in self.ranges.values() for r in rs)

    def stage_done(self, stage):
        if self.world == 1 or stage in self.launched:
            return
        self.launched.add(stage)
        g = self.engine.store.g
        for lo, hi in self.ranges[stage]:
            self.pending.append(dist.all_reduce(g[lo:hi], op=dist.ReduceOp.SUM, group=self.group,
                                                async_op=True))

    def __call__(self, g):
        if self.world > 1:
            for stage in self.ranges:              # anything the backward did not announce
                self.stage_done(stage)
            for w in self.pending:
                w.wait()
        self.pending, self.launched = [], set()
        return 1.0 / self.world


def attach(solver, group=None):
    """Make `solver` data-parallel over the default (or given) process group."""
    sync = GradSync(solver.net.engine, group)
    solver.grad_sync = sync
    solver.net.engine.progress = sync.stage_done
    return sync


def broadcast_params(engine, src=0, group=None):
    """Start all ranks from rank `src`'s parameters."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(engine.store.p, src=src, group=group)
        engine.params_changed()

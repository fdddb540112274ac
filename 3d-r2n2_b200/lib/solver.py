"""Mirror of lib/solver.py: ADAM / SGD update rules and the Solver train / test surface.

The reference compiles `theano.function([x, y], loss, updates=...)`; here one training step is
engine.forward(training) -> engine.backward() -> [NCCL all-reduce of the flat gradient buffer] ->
one fused optimiser kernel over the flat parameter buffer (r2n2_adam / r2n2_sgd).
numpy in / numpy out, like the reference.
"""
import math
import os
import sys
from datetime import datetime

import numpy as np
import torch

from .config import cfg
from .. import ops


def max_or_nan(params, quiet=False):
    """lib/solver.py:12-17"""
    nan_or_max_param = 0.0
    for param_idx, param in enumerate(params):
        nan_or_max_param = np.max(np.abs(param.val.get_value()))
        if not quiet:
            print('param %d : %f' % (param_idx, nan_or_max_param))
    return nan_or_max_param


class Solver(object):

    def __init__(self, net, dist_group=None):
        self.net = net
        self.lr = np.float32(1)
        self.iteration = np.float32(0)     # lib/solver.py:79: float32, starts from 0
        self._policy = None
        self.rank, self.world, self.dist_group = 0, 1, None      # set by dist.attach() (data-parallel training)
        self.grad_sync = None              # set by dist.attach() (bucketed all-reduce of store.g; also sets engine.progress)
        self.compile_model(cfg.TRAIN.POLICY)

    def compile_model(self, policy=cfg.TRAIN.POLICY):
        """lib/solver.py:85-97"""
        if policy not in ('sgd', 'adam'):
            sys.exit('Error: Unimplemented optimization policy')
        self._policy = policy
        self.beta_1, self.beta_2, self.epsilon = 0.9, 0.999, 1e-8   # lib/solver.py:20
        self.w_decay = cfg.TRAIN.WEIGHT_DECAY
        self.momentum = cfg.TRAIN.MOMENTUM

    def set_lr(self, lr):
        self.lr = np.float32(lr)

    # ---- one optimisation step -----------------------------------------------------------
    def _apply_updates_layerwise(self, grad_scale):
        """A user-defined network (models.net.Net define-by-run): one optimiser launch per Weight of
        lib/layers.py, same update rules (lib/solver.py:20-71), L2 decay on non-bias tensors only."""
        from . import layers as ly
        t = float(self.iteration)
        for w in self.net._weights:
            p = w._param
            if p is None:
                continue
            n_decay = 0 if w.is_bias else p.data.numel()
            if w._m is None:
                w._m = torch.zeros_like(p.data)
                w._v = torch.zeros_like(p.data)
            if self._policy == 'adam':
                lr_t = float(self.lr) * math.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t)
                ops.adam(p.data, p.grad, w._m, w._v, p.mirror, n_decay, lr_t, self.beta_1, self.beta_2,
                         self.epsilon, self.w_decay, grad_scale)
            else:
                ops.sgd(p.data, p.grad, w._m, p.mirror, n_decay, float(self.lr), self.momentum, self.w_decay,
                        grad_scale)
            w._host_stale = True
        ly._layer_store.version += 1

    def _apply_updates(self, grad_scale=1.0, ranges=None, finish=True):
        """ranges: element ranges [(lo, hi)] of the flat store to update (64-element aligned segment boundaries; default: all);
        finish=False: more ranges of the same step follow (data-parallel split update, see dist.GradSync.wait_early)."""
        if getattr(self.net, 'engine', None) is None:
            return self._apply_updates_layerwise(grad_scale)
        st = self.net.engine.store
        split = getattr(st, 'mirror_lo', None) is not None           # bf16x3: hi / lo operand halves, one extra pass
        if ranges is None:
            ranges = [(0, st.p.numel())]
        mirror = None if split else st.mirror
        if self._policy == 'adam':
            # lib/solver.py:24-25: lr_t = lr * sqrt(1 - beta_2^t) / (1 - beta_1^t), t float32
            t = float(self.iteration)
            lr_t = float(self.lr) * math.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t)
            st.ensure_moments()
        elif st.m is None:
            st.m = torch.zeros_like(st.p)
        for lo, hi in ranges:
            if hi <= lo:
                continue
            nd = min(max(st.n_decay - lo, 0), hi - lo)               # L2 decay applies to [0, n_decay) of the whole store
            mr = None if mirror is None else mirror[lo:hi]
            if self._policy == 'adam':
                ops.adam(st.p[lo:hi], st.g[lo:hi], st.m[lo:hi], st.v[lo:hi], mr, nd, lr_t, self.beta_1,
                         self.beta_2, self.epsilon, self.w_decay, grad_scale)
            else:
                ops.sgd(st.p[lo:hi], st.g[lo:hi], st.m[lo:hi], mr, nd, float(self.lr), self.momentum,
                        self.w_decay, grad_scale)
        if finish:
            if split:
                st.refresh_mirror()
            st.touch()

    def _step(self, x, y):
        out = self.net.forward_backward(x, y)
        scale = 1.0
        if self.grad_sync is not None and getattr(self.net, 'engine', None) is None:
            raise NotImplementedError('data-parallel training needs the flat gradient store of an engine network')
        gs = self.grad_sync
        if gs is not None and hasattr(gs, 'wait_early') and os.environ.get('R2N2_SPLIT_UPDATE'):
            # OPT-IN experiment: the optimiser pass over everything but the last bucket runs while that bucket's all-reduce
            # is in flight.  Measured on 2 x B200 (ABAB, 30 steps): 10.44 / 10.38 ms with it, 10.33 / 10.34 ms without --
            # three optimiser launches and two extra stream waits cost more than the ~50 us of collective latency they
            # hide, so the default is ONE pass after all buckets have landed.
            scale, early, late = gs.wait_early()
            self._apply_updates(scale, ranges=early, finish=not late)
            if late:
                gs.wait_late()
                self._apply_updates(scale, ranges=late)
            return out['loss']
        if gs is not None:
            scale = gs(self.net.engine.store.g)
        self._apply_updates(scale)
        return out['loss']

    @property
    def train_loss(self):
        """lib/solver.py:102-109: a property returning the compiled step; the iteration counter is
        bumped on ACCESS, before the call (first step has t = 1)."""
        self.iteration = np.float32(self.iteration + 1)

        def fn(x, y):
            loss = float(self._step(x, y))       # host sync: the step's kernels have finished
            self._check_device_errors()
            return loss
        return fn

    @staticmethod
    def _check_device_errors():
        """A tcgen05 pipeline wait that timed out leaves a sticky per-device error word (csrc/umma.cuh
        mbar_wait) and garbage accumulators: surface it wherever the host synchronises anyway."""
        if ops.umma_error_flag() != 0:
            from .._lib import R2N2Error
            raise R2N2Error('a tcgen05/TMA pipeline timed out on the device: results of this step are invalid')

    def train_step_device(self, x_dev, y_dev):
        """Same step with inputs already on the device; returns the loss as a 0-d CUDA tensor
        (no host sync).  Used by bench.py for the HBM-resident `value` measurement."""
        self.iteration = np.float32(self.iteration + 1)
        return self._step(x_dev, y_dev)

    def train(self, train_queue, val_queue=None):
        ''' Given data queues, train the network (lib/solver.py:111-186) '''
        save_dir = os.path.join(cfg.DIR.OUT_PATH)
        main_rank = self.rank == 0             # data-parallel: every rank steps, rank 0 prints and saves
        if main_rank and not os.path.exists(save_dir):
            os.makedirs(save_dir, exist_ok=True)
        training_losses = []
        start_iter = 0
        if cfg.TRAIN.RESUME_TRAIN:
            self.net.load(cfg.CONST.WEIGHTS)
            start_iter = cfg.TRAIN.INITIAL_ITERATION
        lr = cfg.TRAIN.DEFAULT_LEARNING_RATE
        lr_steps = [int(k) for k in cfg.TRAIN.LEARNING_RATES.keys()]
        if main_rank:
            print('Set the learning rate to %f.' % lr)
        self.set_lr(lr)
        for train_ind in range(start_iter, cfg.TRAIN.NUM_ITERATION + 1):
            batch_img, batch_voxel = train_queue.get()
            if self.net.is_x_tensor4:
                batch_img = batch_img[0]
            loss = self.train_loss(batch_img, batch_voxel)
            training_losses.append(loss)
            if train_ind in lr_steps:
                self.set_lr(float(cfg.TRAIN.LEARNING_RATES[str(train_ind)]))
                if main_rank:
                    print('Learing rate decreased to %f: ' % self.lr)
            if train_ind % cfg.TRAIN.PRINT_FREQ == 0 and main_rank:
                print('%s Iter: %d Loss: %f' % (datetime.now(), train_ind, loss))
            if train_ind % cfg.TRAIN.VALIDATION_FREQ == 0 and val_queue is not None:
                val_losses = []
                for i in range(cfg.TRAIN.NUM_VALIDATION_ITERATIONS):
                    batch_img, batch_voxel = val_queue.get()
                    _, val_loss, _ = self.test_output(batch_img, batch_voxel)
                    val_losses.append(val_loss)
                val_mean = float(np.mean(val_losses))
                if self.world > 1:             # every rank validated its own shard
                    from ..dist import all_reduce_mean_scalar
                    val_mean = all_reduce_mean_scalar(val_mean, self.dist_group)
                if main_rank:
                    print('%s Test loss: %f' % (datetime.now(), val_mean))
            if train_ind % cfg.TRAIN.NAN_CHECK_FREQ == 0:
                # (parameters are bit-identical on every rank: all ranks reach the same verdict)
                max_param = max_or_nan(self.net.params, quiet=not main_rank)
                if np.isnan(max_param):
                    print('NAN detected')
                    break
            if train_ind % cfg.TRAIN.SAVE_FREQ == 0 and not train_ind == 0 and main_rank:
                self.save(training_losses, save_dir, train_ind)
            if self.world > 1:
                # ranks see different batches: stop together, on the mean loss, or the next all-reduce would hang
                from ..dist import all_reduce_mean_scalar
                stop_loss = all_reduce_mean_scalar(loss, self.dist_group)
            else:
                stop_loss = loss
            if stop_loss > cfg.TRAIN.LOSS_LIMIT:
                print("Cost exceeds the threshold. Stop training")
                break
        return training_losses

    def save(self, training_losses, save_dir, step):
        ''' lib/solver.py:188-205 '''
        save_path = os.path.join(save_dir, 'weights.%d' % (step))
        self.net.save(save_path)
        symlink_path = os.path.join(save_dir, 'weights.npy')
        if os.path.lexists(symlink_path):
            os.remove(symlink_path)
        os.symlink("%s.npy" % os.path.abspath(save_path), symlink_path)
        with open(os.path.join(save_dir, 'loss.%d.txt' % step), 'w') as f:
            f.write('\n'.join([str(l) for l in training_losses]))

    def test_output(self, x, y=None):
        '''lib/solver.py:207-240: returns (prediction, activations) or
        (prediction, loss, activations), all numpy.'''
        out = self.net.forward(x, y)
        prediction = out['pred'].cpu().numpy()
        activations = [out['update_gates'].float().contiguous().cpu().numpy()]
        self._check_device_errors()
        if y is None:
            return prediction, activations
        return prediction, float(out['loss']), activations

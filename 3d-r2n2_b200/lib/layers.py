"""Mirror of the reference's layer library (lib/layers.py) on top of libr2n2_b200.so.

Same class names, constructor signatures, `.output_shape` / `.output` / `.params` protocol and
parameter sharing via `params=` as the reference, so a `network_definition` body written
against the reference runs unchanged except that `theano.scan` becomes a host loop
(SURVEY.md §8b).  Differences that are inherent to leaving Theano:
  * `.output` is evaluated eagerly on first access (a CUDA tensor exposed with the REFERENCE
    logical shape, e.g. (N,C,H,W) / (B,D,C,H,W), as a stride-permuted view of the channels-last
    buffer the kernels use) instead of being a symbolic variable;
  * `Weight.val` is a small shared-variable shim (`get_value` / `set_value`, reference layout).
Every `set_output` cites the reference lines it replaces.  There is no CPU fallback.
"""
import collections.abc

import numpy as np
import torch

from .. import _lib as L
from .. import ops
from .. import packing as pk
from ..ops import Act, Param, Ctx

trainable_params = []          # lib/layers.py:10 (module-global, never cleared: quirk 13)
_rng = np.random.RandomState()  # lib/layers.py:31 uses an unseeded RandomState per Weight
_ctx = None
DEVICE = 'cuda'              # the kernels only run on CUDA; tests of the host logic override it
_replay = None               # active replay_weights context (define-by-run nets, models/net.py)


class _LayerStore:
    """What ops.Param needs from a parameter store: a version that moves when the optimiser (or set_value)
    changes the values, so that transposed data-gradient operands are rebuilt."""
    version = 0


_layer_store = _LayerStore()


class replay_weights:
    """A user-defined `network_definition` (models/net.py:43-51) runs once per call here instead of once per
    Theano compile.  Inside this context every `Weight(...)` construction returns the tensor created at the same
    position of the FIRST run instead of a new one, so the body needs no change."""

    def __init__(self, weights):
        self.weights, self.pos = list(weights), 0

    def __enter__(self):
        global _replay
        if _replay is not None:
            raise RuntimeError('nested network_definition replay')
        _replay = self
        self.pos = 0
        return self

    def __exit__(self, et, ev, tb):
        global _replay
        _replay = None
        if et is None and self.pos != len(self.weights):
            raise RuntimeError('network_definition created %d Weights, %d on the first run'
                               % (self.pos, len(self.weights)))
        return False

    def next(self, w_shape, is_bias):
        if self.pos >= len(self.weights):
            raise RuntimeError('network_definition creates more Weights than on its first run')
        w = self.weights[self.pos]
        self.pos += 1
        if tuple(np.atleast_1d(w.shape)) != tuple(np.atleast_1d(w_shape)) or w.is_bias != is_bias:
            raise RuntimeError('network_definition is not deterministic: Weight %d was %s, now %s'
                               % (self.pos - 1, tuple(np.atleast_1d(w.shape)), tuple(np.atleast_1d(w_shape))))
        return w


def get_trainable_params():
    """lib/layers.py:13-15"""
    global trainable_params
    return trainable_params


def seed_weights(seed):
    """Not in the reference (its init is unseeded): make Weight init reproducible."""
    global _rng
    _rng = np.random.RandomState(seed)


def set_context(ctx):
    """Execution context (compute dtype / conv algorithm) used by layers built afterwards."""
    global _ctx
    _ctx = ctx


def get_context():
    global _ctx
    if _ctx is None:
        _ctx = Ctx(torch.float32, L.ALGO_SIMT, training=False)
    return _ctx


class _Shared:
    """theano.shared stand-in: host master copy in the reference layout."""

    def __init__(self, owner):
        self._owner = owner

    def get_value(self):
        self._owner.sync_host()
        return self._owner.np_values

    def set_value(self, v):
        v = np.asarray(v, dtype=np.float32)
        if tuple(v.shape) != tuple(np.shape(self._owner.np_values)):
            raise ValueError('shape mismatch: %s vs %s' % (v.shape, np.shape(self._owner.np_values)))
        self._owner.np_values = v
        self._owner._host_stale = False
        self._owner.repack()


class Weight(object):
    """lib/layers.py:18-79: same signature, same fillers, same fan-in/out rules."""

    def __new__(cls, w_shape, is_bias, *args, **kwargs):
        if _replay is not None:
            return _replay.next(w_shape, is_bias)
        return super(Weight, cls).__new__(cls)

    def __init__(self, w_shape, is_bias, mean=0, std=0.01, filler='msra', fan_in=None,
                 fan_out=None, name=None):
        if getattr(self, '_defined', False):      # handed out again by replay_weights
            return
        super(Weight, self).__init__()
        assert (is_bias in [True, False])
        rng = _rng
        if isinstance(w_shape, collections.abc.Iterable) and not is_bias:
            if len(w_shape) > 1 and len(w_shape) < 5:
                fan_in = np.prod(w_shape[1:])
                fan_out = np.prod(w_shape) / w_shape[1]
                n = (fan_in + fan_out) / 2.
            elif len(w_shape) == 5:
                fan_in = np.prod(w_shape[1:])
                fan_out = np.prod(w_shape) / w_shape[2]
                n = (fan_in + fan_out) / 2.
            else:
                raise NotImplementedError(
                    'Filter shape with ndim > 5 not supported: len(w_shape) = %d' % len(w_shape))
        else:
            n = 1
        if fan_in and fan_out:
            n = (fan_in + fan_out) / 2.
        if filler == 'gaussian':
            self.np_values = np.asarray(rng.normal(mean, std, w_shape), dtype=np.float32)
        elif filler == 'msra':
            self.np_values = np.asarray(rng.normal(mean, np.sqrt(2. / n), w_shape), dtype=np.float32)
        elif filler == 'xavier':
            scale = np.sqrt(3. / n)
            self.np_values = np.asarray(rng.uniform(low=-scale, high=scale, size=w_shape),
                                        dtype=np.float32)
        elif filler == 'constant':
            self.np_values = (mean * np.ones(w_shape, dtype=np.float32)).astype(np.float32)
        elif filler == 'orth':
            ndim = np.prod(w_shape)
            W = np.random.randn(ndim, ndim)
            u, s, v = np.linalg.svd(W)
            self.np_values = u.astype(np.float32).reshape(w_shape)
        else:
            raise NotImplementedError('Filler %s not implemented' % filler)
        self.is_bias = is_bias
        self.val = _Shared(self)
        self.shape = w_shape
        self.name = name
        self._kind = None      # how the owning layer packs it: ('conv2d', cin_pad) ...
        self._param = None
        self._host_stale = False   # the optimiser moved the device copy: np_values is refreshed on get_value
        self._m = self._v = None   # optimiser moments (packed layout, like the parameter)
        self.grad_seen = False
        self._defined = True
        global trainable_params
        trainable_params.append(self)

    # ---- packed device operand (built lazily, rebuilt after set_value) -----------------
    def bind(self, kind):
        if self._kind is None:
            self._kind = kind

    def repack(self):
        """Values changed on the host: refill the packed device copy IN PLACE (optimiser moments and the
        gradient buffer stay attached) or drop it when the layout is not known yet."""
        if self._param is None:
            return
        fresh = self._pack(torch.as_tensor(self.np_values).to(self._param.data.device))[0]
        self._param.data.copy_(fresh.contiguous().view(-1))
        if self._param.mirror is not None:
            ops.cast(self._param.data, self._param.mirror)
        _layer_store.version += 1

    def sync_host(self):
        """Device copy -> np_values in the reference layout (after optimiser steps)."""
        if not self._host_stale or self._param is None:
            return
        self.np_values = self._unpack(self._param.data)
        self._host_stale = False

    def grad_value(self):
        """tensor.grad(loss, self.val) of the last forward_backward, reference layout (models/net.py:58)."""
        if self._param is None:
            return np.zeros(tuple(np.atleast_1d(self.shape)), np.float32)
        return self._unpack(self._param.grad)

    def _unpack(self, v):
        kind, shp = self._kind, tuple(int(d) for d in np.atleast_1d(self.shape))
        if kind is None or kind[0] == 'bias':
            r = v.view(shp)
        elif kind[0] == 'conv2d':
            r = pk.unpack_conv2d(v.view(shp[0], shp[2], shp[3], kind[1]), shp[1])
        elif kind[0] == 'conv3d':
            r = pk.unpack_conv3d(v.view(shp[0], shp[1], shp[3], shp[4], shp[2]))
        elif kind[0] == 'fc':
            r = pk.unpack_fc(v.view(shp[1], shp[0]), kind[1])
        elif kind[0] == 'wx':
            d, c, h, w = kind[1]
            r = pk.unpack_fcconv_wx(v.view(d * h * w, c, shp[0]), d, c, h, w)
        elif kind[0] == 'blockdiag':
            r = v.view(shp[:-2] + (shp[-1], shp[-2])).transpose(-1, -2)
        else:
            raise ValueError(kind)
        return r.contiguous().cpu().numpy().reshape(shp)

    def _pack(self, t):
        kind = self._kind
        if kind is None or kind[0] == 'bias':
            p, cout, taps, cin = t.contiguous(), t.numel(), 1, 1
        elif kind[0] == 'conv2d':
            p = pk.pack_conv2d(t, kind[1])
            cout, taps, cin = p.shape[0], p.shape[1] * p.shape[2], p.shape[3]
        elif kind[0] == 'conv3d':
            p = pk.pack_conv3d(t)
            cout, taps, cin = p.shape[0], p.shape[1] * p.shape[2] * p.shape[3], p.shape[4]
        elif kind[0] == 'fc':
            p = pk.pack_fc(t, kind[1])
            cout, taps, cin = p.shape[0], 1, p.shape[1]
        elif kind[0] == 'wx':
            d, c, h, w = kind[1]
            p = pk.pack_fcconv_wx(t, d, c, h, w).contiguous()
            cout, taps, cin = d * h * w * c, 1, t.shape[0]
        elif kind[0] == 'blockdiag':          # (D_1..D_{n-1}, n_in, n_out) -> per position [n_out][n_in]
            p = t.transpose(-1, -2).contiguous()
            cout, taps, cin = t.shape[-1], 1, t.shape[-2]
        else:
            raise ValueError(kind)
        return p, cout, taps, cin

    def param(self, ctx):
        if self._param is not None and self._param_dtype == ctx.dtype:
            return self._param
        self.sync_host()
        p, cout, taps, cin = self._pack(torch.as_tensor(self.np_values).to(DEVICE))
        p = p.contiguous().view(-1)
        mirror = None
        if ctx.dtype != torch.float32 and not self.is_bias:
            mirror = torch.empty_like(p, dtype=ctx.dtype)
            ops.cast(p, mirror)
        self._param = Param(p, torch.zeros_like(p), mirror, cout, taps, cin, _layer_store)
        self._param_dtype = ctx.dtype
        self._m = self._v = None
        return self._param


def _cin_pad(ctx, cin):
    if cin % 4 == 0 and (ctx.algo == L.ALGO_SIMT or cin % 16 == 0):
        return cin
    return (cin + 3) // 4 * 4 if ctx.algo == L.ALGO_SIMT else (cin + 15) // 16 * 16


def _as_act(t, shape, ctx):
    """Reference-layout tensor / ndarray (or an Act) -> channels-last Act."""
    if isinstance(t, Act):
        return t
    if getattr(t, '_r2n2_act', None) is not None:      # a Layer's .output: stay on the tape
        return t._r2n2_act
    if isinstance(t, np.ndarray):
        t = torch.as_tensor(t)
    if not isinstance(t, torch.Tensor):
        raise ValueError('InputLayer needs a numpy array or torch tensor, got %r' % type(t))
    t = t.to(DEVICE)
    if tuple(t.shape) != tuple(shape):
        raise ValueError('InputLayer: tensor shape %s does not match declared shape %s'
                         % (tuple(t.shape), tuple(shape)))
    if t.dim() == 4:       # (N,C,H,W)
        N, C, H, W = t.shape
        cp = _cin_pad(ctx, C)
        y = torch.empty(N * H * W * cp, dtype=ctx.dtype, device=t.device)
        ops.nchw_to_nhwc(t.float().contiguous(), y, cp)
        return Act(y, (N, 1, H, W, cp))
    if t.dim() == 5:       # (B,D,C,H,W)
        B, D, C, H, W = t.shape
        y = pk.to_channels_last_3d(t).to(ctx.dtype).view(-1)
        return Act(y, (B, D, H, W, C))
    if t.dim() == 2:
        return Act(t.to(ctx.dtype).contiguous().view(-1), (t.shape[0], 1, 1, 1, t.shape[1]))
    if t.dim() == 3:
        return Act(t.to(ctx.dtype).contiguous().view(-1), (t.shape[0], 1, 1, t.shape[1], t.shape[2]))
    raise ValueError('unsupported input rank %d' % t.dim())


def _ref_view(act, shape):
    """channels-last Act -> tensor with the reference's logical shape."""
    N, D, H, W, C = act.geom
    d = act.data.view(N, D, H, W, C)
    if len(shape) == 4:
        v = d[:, 0].permute(0, 3, 1, 2)[:, :shape[1]]
    elif len(shape) == 5:
        v = d.permute(0, 1, 4, 2, 3)
    elif len(shape) == 2:
        v = d.reshape(N, -1)
    elif len(shape) == 3:
        v = d.reshape(N, shape[1], shape[2])
    else:
        raise ValueError(shape)
    if v.data_ptr() == act.data.data_ptr() and v.numel() == act.data.numel():
        # the symbolic variable of the reference: feeding it to InputLayer / SoftmaxWithLoss3D / scan keeps the
        # producing Act (and with it the gradient path), like passing a Theano variable does
        v._r2n2_act = act
    return v


class InputLayer(object):
    """lib/layers.py:82-97"""

    def __init__(self, input_shape, tinput=None):
        self._output_shape = input_shape
        self._input = tinput
        self._act = None

    @property
    def act(self):
        if self._input is None:
            raise ValueError('Cannot call output for the layer. Initialize' \
                             + ' the layer with an input argument')
        if self._act is None:
            self._act = _as_act(self._input, self._output_shape, get_context())
        return self._act

    @property
    def output(self):
        if self._input is None:
            raise ValueError('Cannot call output for the layer. Initialize' \
                             + ' the layer with an input argument')
        return _ref_view(self.act, self._output_shape)

    @property
    def output_shape(self):
        return self._output_shape


class Layer(object):
    """lib/layers.py:100-130"""

    def __init__(self, prev_layer):
        self._output = None
        self._act = None
        self._output_shape = None
        self._prev_layer = prev_layer
        self._input_shape = prev_layer.output_shape

    def set_output(self):
        raise NotImplementedError('Layer virtual class')

    @property
    def output_shape(self):
        if self._output_shape is None:
            raise ValueError('Set output shape first')
        return self._output_shape

    @property
    def act(self):
        if self._act is None:
            self.set_output()
        return self._act

    @property
    def output(self):
        shape = self._output_shape if self._output_shape is not None else self._input_shape
        return _ref_view(self.act, shape)


class TensorProductLayer(Layer):
    """lib/layers.py:133-161"""

    def __init__(self, prev_layer, n_out, params=None, bias=True):
        super().__init__(prev_layer)
        self._bias = bias
        n_in = self._input_shape[-1]
        if params is None:
            self.W = Weight((n_in, n_out), is_bias=False)
            if bias:
                self.b = Weight((n_out,), is_bias=True, mean=0.1, filler='constant')
        else:
            self.W = params[0]
            if bias:
                self.b = params[1]
        self.W.bind(('fc', getattr(prev_layer, '_chw', None)))
        self.params = [self.W]
        if bias:
            self.params.append(self.b)
        self._output_shape = [self._input_shape[0]]
        self._output_shape.extend(self._input_shape[1:-1])
        self._output_shape.append(n_out)

    def set_output(self):
        # tensor.dot(prev.output, W) + b      (lib/layers.py:159-161)
        ctx = get_context()
        x = self._prev_layer.act
        N = x.geom[0]
        xin = ctx.view(x, (N, 1, 1, 1, x.data.numel() // N))
        self._act = ctx.conv(xin, self.W.param(ctx), self.b.param(ctx) if self._bias else None,
                             (1, 1, 1))


class AddLayer(Layer):
    """lib/layers.py:206-214"""

    def __init__(self, prev_layer, add_layer):
        super().__init__(prev_layer)
        self._output_shape = self._input_shape
        self._add_layer = add_layer

    def set_output(self):
        self._act = get_context().add_act(self._prev_layer.act, self._add_layer.act)


class EltwiseMultiplyLayer(Layer):
    """lib/layers.py:217-225"""

    def __init__(self, prev_layer, mult_layer):
        super().__init__(prev_layer)
        self._output_shape = self._input_shape
        self._mult_layer = mult_layer

    def set_output(self):
        self._act = get_context().mul(self._prev_layer.act, self._mult_layer.act)


class FlattenLayer(Layer):
    """lib/layers.py:228-236: flatten(2) is (C,H,W)-major in the reference; here the buffer stays
    channels-last and the consumer's weight rows are permuted at pack time (`_chw`)."""

    def __init__(self, prev_layer):
        super().__init__(prev_layer)
        self._output_shape = [self._input_shape[0], int(np.prod(self._input_shape[1:]))]
        self._chw = tuple(int(s) for s in self._input_shape[1:]) if len(self._input_shape) == 4 else None

    def set_output(self):
        x = self._prev_layer.act
        self._act = get_context().view(x, (x.geom[0], 1, 1, 1, x.data.numel() // x.geom[0]))

    @property
    def output(self):
        # present the reference's (C,H,W)-major ordering
        prev = self._prev_layer.output
        return prev.reshape(prev.shape[0], -1)


class ConvLayer(Layer):
    """lib/layers.py:266-333.  filter_shape: [n_out_channel, n_height, n_width]"""

    def __init__(self, prev_layer, filter_shape, padding=True, params=None):
        super().__init__(prev_layer)
        self._padding = padding
        self._filter_shape = [filter_shape[0], self._input_shape[1], filter_shape[1],
                              filter_shape[2]]
        if params is None:
            self.W = Weight(self._filter_shape, is_bias=False)
            self.b = Weight((filter_shape[0],), is_bias=True, mean=0.1, filler='constant')
        else:
            for i, s in enumerate(self._filter_shape):
                assert (params[0].shape[i] == s)
            self.W = params[0]
            self.b = params[1]
        self.params = [self.W, self.b]
        if padding and filter_shape[1] * filter_shape[2] > 1:
            self._padding = [0, 0, int((filter_shape[1] - 1) / 2), int((filter_shape[2] - 1) / 2)]
            self._output_shape = [self._input_shape[0], filter_shape[0], self._input_shape[2],
                                  self._input_shape[3]]
        else:
            self._padding = [0] * 4
            self._output_shape = [self._input_shape[0], filter_shape[0],
                                  self._input_shape[2] - filter_shape[1] + 1,
                                  self._input_shape[3] - filter_shape[2] + 1]
        if self._output_shape[2] != self._input_shape[2] or self._output_shape[3] != self._input_shape[3]:
            raise NotImplementedError("only 'same' convolutions are on the accelerated path "
                                      '(every ConvLayer of the shipped models is)')

    def set_output(self):
        # zero pad + tensor.nnet.conv2d(valid, flipped kernel) + b   (lib/layers.py:304-333)
        ctx = get_context()
        x = self._prev_layer.act
        self.W.bind(('conv2d', x.geom[4]))
        k = (1, self._filter_shape[2], self._filter_shape[3])
        self._act = ctx.conv(x, self.W.param(ctx), self.b.param(ctx), k)


class PoolLayer(Layer):
    """lib/layers.py:336-354"""

    def __init__(self, prev_layer, pool_size=(2, 2), padding=(1, 1)):
        super().__init__(prev_layer)
        self._pool_size = pool_size
        self._padding = padding
        if tuple(pool_size) != (2, 2) or tuple(padding) != (1, 1):
            raise NotImplementedError('only pool_size=(2,2), padding=(1,1) is accelerated')
        img_rows = self._input_shape[2] + 2 * padding[0]
        img_cols = self._input_shape[3] + 2 * padding[1]
        out_r = (img_rows - pool_size[0]) // pool_size[0] + 1
        out_c = (img_cols - pool_size[1]) // pool_size[1] + 1
        self._output_shape = [self._input_shape[0], self._input_shape[1], out_r, out_c]

    def set_output(self):
        # pool.pool_2d(ds=(2,2), ignore_border=True, padding=(1,1))   (lib/layers.py:349-353)
        self._act = get_context().pool(self._prev_layer.act)


class Unpool3DLayer(Layer):
    """lib/layers.py:357-386"""

    def __init__(self, prev_layer, unpool_size=(2, 2, 2), padding=(0, 0, 0)):
        super().__init__(prev_layer)
        self._unpool_size = unpool_size
        self._padding = padding
        if tuple(unpool_size) != (2, 2, 2) or tuple(padding) != (0, 0, 0):
            raise NotImplementedError('only unpool_size=(2,2,2), padding=(0,0,0) is accelerated')
        self._output_shape = (self._input_shape[0],
                              unpool_size[0] * self._input_shape[1] + 2 * padding[0],
                              self._input_shape[2],
                              unpool_size[1] * self._input_shape[3] + 2 * padding[1],
                              unpool_size[2] * self._input_shape[4] + 2 * padding[2])

    def set_output(self):
        # alloc(0) + set_subtensor(out[:, ::2, :, ::2, ::2], x)   (lib/layers.py:375-385)
        self._act = get_context().unpool(self._prev_layer.act)


class Conv3DLayer(Layer):
    """lib/layers.py:389-442"""

    def __init__(self, prev_layer, filter_shape, padding=None, params=None):
        super().__init__(prev_layer)
        self._filter_shape = [filter_shape[0], filter_shape[1], self._input_shape[2],
                              filter_shape[2], filter_shape[3]]
        self._padding = padding
        if params is None:
            self.W = Weight(self._filter_shape, is_bias=False)
            self.b = Weight((filter_shape[0],), is_bias=True, mean=0.1, filler='constant')
            params = [self.W, self.b]
        else:
            self.W = params[0]
            self.b = params[1]
        self.W.bind(('conv3d',))
        self.params = [self.W, self.b]
        if padding is None:
            self._padding = [0, int((filter_shape[1] - 1) / 2), 0, int((filter_shape[2] - 1) / 2),
                             int((filter_shape[3] - 1) / 2)]
        else:
            raise NotImplementedError("only the default 'same' padding is accelerated")
        self._output_shape = [self._input_shape[0], self._input_shape[1], filter_shape[0],
                              self._input_shape[3], self._input_shape[4]]

    def set_output(self):
        # zero pad + conv3d2d.conv3d + b   (lib/layers.py:424-442)
        ctx = get_context()
        fs = self._filter_shape
        self._act = ctx.conv(self._prev_layer.act, self.W.param(ctx), self.b.param(ctx),
                             (fs[1], fs[3], fs[4]))


class FCConv3DLayer(Layer):
    """lib/layers.py:445-505: 3D convolution of the hidden grid + FC projection of the input"""

    def __init__(self, prev_layer, fc_layer, filter_shape, padding=None, params=None):
        super().__init__(prev_layer)
        self._fc_layer = fc_layer
        self._filter_shape = [filter_shape[0], filter_shape[2], filter_shape[1], filter_shape[3],
                              filter_shape[4]]
        self._padding = padding
        if padding is None:
            self._padding = [0, int((self._filter_shape[1] - 1) / 2), 0,
                             int((self._filter_shape[3] - 1) / 2), int((self._filter_shape[4] - 1) / 2)]
        else:
            raise NotImplementedError("only the default 'same' padding is accelerated")
        self._output_shape = [self._input_shape[0], self._input_shape[1], filter_shape[0],
                              self._input_shape[3], self._input_shape[4]]
        if params is None:
            self.Wh = Weight(self._filter_shape, is_bias=False)
            self._Wx_shape = [self._fc_layer._output_shape[1], int(np.prod(self._output_shape[1:]))]
            self.Wx = Weight(self._Wx_shape, is_bias=False, fan_in=self._input_shape[1],
                             fan_out=self._output_shape[2])
            self.b = Weight((filter_shape[0],), is_bias=True, mean=0.1, filler='constant')
            params = [self.Wh, self.Wx, self.b]
        else:
            self.Wh = params[0]
            self.Wx = params[1]
            self.b = params[2]
        self.Wh.bind(('conv3d',))
        os_ = self._output_shape
        self.Wx.bind(('wx', (os_[1], os_[2], os_[3], os_[4])))
        self.params = [self.Wh, self.Wx, self.b]

    def set_output(self):
        # conv3d2d.conv3d(pad(h), Wh) + reshape(dot(fc, Wx)) + b   (lib/layers.py:489-505)
        ctx = get_context()
        h = self._prev_layer.act
        f = self._fc_layer.act
        B = f.geom[0]
        fin = ctx.view(f, (B, 1, 1, 1, f.data.numel() // B))
        xp = ctx.conv(fin, self.Wx.param(ctx), None, (1, 1, 1))
        xp = ctx.view(xp, (h.geom[0], h.geom[1], h.geom[2], h.geom[3], self._output_shape[2]))
        fs = self._filter_shape
        self._act = ctx.conv(h, self.Wh.param(ctx), self.b.param(ctx), (fs[1], fs[3], fs[4]),
                             residual=xp)


class SoftmaxWithLoss3D(object):
    """lib/layers.py:578-601.  Softmax with loss (n_batch, n_vox, n_label, n_vox, n_vox)"""

    def __init__(self, input):
        # `input` is a Layer's .output in the reference; accept the layer, its Act or a
        # reference-layout tensor
        ctx = get_context()
        if isinstance(input, Layer):
            input = input.act
        if not isinstance(input, Act):
            t = input
            input = _as_act(t, tuple(t.shape), Ctx(torch.float32, ctx.algo))
        self._src = input               # the Act on the tape: dL/dlogits is delivered here by loss()
        if input.data.dtype != torch.float32:
            input = Act(input.data.float(), input.geom)
        self.input = input
        B, D, H, W, C = input.geom
        if C != 2 or not (D == H == W):
            raise ValueError('SoftmaxWithLoss3D is accelerated for (B,n,2,n,n) inputs only')
        self._B, self._nv = B, D

    def _run(self, y, want_pred, want_grad=False):
        B, nv = self._B, self._nv
        dev = self.input.data.device
        pred = torch.empty(B, nv, 2, nv, nv, dtype=torch.float32, device=dev) if want_pred else None
        loss = None
        yt = None
        if y is not None:
            yt = torch.as_tensor(y).to(dev).float().contiguous()
            loss = torch.zeros(1, dtype=torch.float32, device=dev)
        dlogits = None
        src = self._src
        if want_grad and y is not None and src.requires_grad and src.n_recv == 0:
            # tensor.grad(loss, ...) starts here (models/net.py:56-58): d(mean CE)/dlogits from the same pass
            dlogits = torch.empty(src.data.numel(), dtype=src.data.dtype, device=dev)
        ops.softmax_ce3d(self.input.data, yt, pred, loss, dlogits, B, nv, 1.0 / float(B * nv ** 3))
        if dlogits is not None:
            src.grad, src.n_recv = dlogits, 1
        return pred, (None if loss is None else loss[0] / float(B * nv ** 3))

    def prediction(self):
        return self._run(None, True)[0]

    def error(self, y, threshold=0.5):
        p = self.prediction()
        yt = torch.as_tensor(y).to(p.device).float()
        return torch.mean(((p >= threshold).float() == yt).float())

    def loss(self, y):
        return self._run(y, False, want_grad=get_context().training)[1]


class LeakyReLU(Layer):
    """lib/layers.py:628-646"""

    def __init__(self, prev_layer, leakiness=0.01):
        super().__init__(prev_layer)
        self._leakiness = leakiness
        self._output_shape = self._input_shape
        if abs(leakiness - 0.01) > 1e-12:
            raise NotImplementedError('the kernels hard-wire leakiness=0.01 (all shipped models)')

    def set_output(self):
        self._act = get_context().lrelu(self._prev_layer.act)


class _Elt(Layer):
    _op = None        # name of the tape-aware Ctx op (forward kernel + its derivative)

    def __init__(self, prev_layer):
        super().__init__(prev_layer)
        self._output_shape = self._input_shape

    def set_output(self):
        self._act = getattr(get_context(), self._op)(self._prev_layer.act)


class SigmoidLayer(_Elt):
    """lib/layers.py:649-656"""
    _op = 'sigmoid'


class TanhLayer(_Elt):
    """lib/layers.py:659-665 (the reference forgets to set _output_shape there)"""
    _op = 'tanh'


class ComplementLayer(_Elt):
    """lib/layers.py:668-676: 1 - input"""
    _op = 'complement'


# ------------------------------------------------------------------------------------------------------
# Layer classes no shipped model instantiates (lib/layers.py:164-203, 239-263, 508-575, 604-625); kept so that
# `from lib.layers import ...` of a user's network file resolves and trains.  Pure data movement (dimshuffle,
# reshape, concatenate) is torch plumbing between channels-last buffers; the arithmetic is C-ABI calls.
def _geom_of(shape):
    """channels-last geometry (N,D,H,W,C) that stores a reference-layout tensor of this shape"""
    s = [int(v) for v in shape]
    if len(s) == 2:
        return (s[0], 1, 1, 1, s[1])
    if len(s) == 3:
        return (s[0], 1, 1, s[1], s[2])
    if len(s) == 4:
        return (s[0], 1, s[2], s[3], s[1])
    if len(s) == 5:
        return (s[0], s[1], s[3], s[4], s[2])
    raise ValueError('unsupported rank %d' % len(s))


def _logical(data, shape):
    """channels-last buffer -> tensor with the reference's logical shape (a view)"""
    g = _geom_of(shape)
    d = data.view(g)
    if len(shape) == 2:
        return d.view(g[0], g[4])
    if len(shape) == 3:
        return d.view(g[0], g[3], g[4])
    if len(shape) == 4:
        return d[:, 0].permute(0, 3, 1, 2)
    return d.permute(0, 1, 4, 2, 3)


def _physical(t, shape):
    """reference-layout tensor -> contiguous channels-last buffer (flat)"""
    if len(shape) == 4:
        t = t.permute(0, 2, 3, 1)
    elif len(shape) == 5:
        t = t.permute(0, 1, 3, 4, 2)
    return t.contiguous().view(-1)


def _movement(ctx, layers, fn, out_shape):
    """out = fn(*[layer.output]) for a pure data-movement fn; the gradient is the same movement backwards."""
    acts = [l.act for l in layers]
    shapes = [tuple(l.output_shape) for l in layers]
    for a, s in zip(acts, shapes):
        if a.data.numel() != int(np.prod(s)):
            raise NotImplementedError('data-movement layers need unpadded channel counts')
    ins = [_logical(a.data, s) for a, s in zip(acts, shapes)]
    out = Act(_physical(fn(*ins), out_shape), _geom_of(out_shape),
              requires_grad=ctx.training and any(a.requires_grad for a in acts))
    if out.requires_grad:
        for a in acts:
            ctx.consume(a)

        def bwd():
            if out.grad is None:
                return
            with torch.enable_grad():
                leaves = [torch.zeros_like(i, dtype=torch.float32).requires_grad_(True) for i in ins]
                gs = torch.autograd.grad(fn(*leaves), leaves, _logical(out.grad, out_shape).float())
            for a, s, g in zip(acts, shapes, gs):
                if a.requires_grad:
                    ctx.accumulate(a, _physical(g, s).to(a.data.dtype))
            out.grad = None
        ctx.tape.append(bwd)
    return out


class DimShuffleLayer(Layer):
    """lib/layers.py:239-250"""

    def __init__(self, prev_layer, shuffle_pattern):
        super().__init__(prev_layer)
        self._shuffle_pattern = shuffle_pattern
        self._output_shape = list(self._input_shape)
        for out_dim, in_dim in enumerate(shuffle_pattern):
            self._output_shape[out_dim] = self._input_shape[in_dim]
        self._output_shape = tuple(self._output_shape)

    def set_output(self):
        pat = tuple(self._shuffle_pattern)
        self._act = _movement(get_context(), [self._prev_layer], lambda t: t.permute(pat), self._output_shape)

    @property
    def output(self):
        return _logical(self.act.data, self._output_shape)


class ReshapeLayer(Layer):
    """lib/layers.py:253-263"""

    def __init__(self, prev_layer, reshape):
        super().__init__(prev_layer)
        self._output_shape = [self._prev_layer.output_shape[0]]
        self._output_shape.extend(reshape)
        self._output_shape = tuple(self._output_shape)
        print('Reshape the prev layer to [%s]' % ','.join(str(x) for x in self._output_shape))

    def set_output(self):
        shp = tuple(int(v) for v in self._output_shape)
        self._act = _movement(get_context(), [self._prev_layer], lambda t: t.reshape(shp), shp)

    @property
    def output(self):
        return _logical(self.act.data, self._output_shape)


class ConcatLayer(Layer):
    """lib/layers.py:604-625"""

    def __init__(self, prev_layers, axis=1):
        assert (len(prev_layers) > 1)
        super().__init__(prev_layers[0])
        self._axis = axis
        self._prev_layers = prev_layers
        self._output_shape = list(self._input_shape).copy()
        for prev_layer in prev_layers[1:]:
            self._output_shape[axis] += prev_layer._output_shape[axis]
        print('Concat the prev layer to [%s]' % ','.join(str(x) for x in self._output_shape))

    def set_output(self):
        ax = self._axis
        self._act = _movement(get_context(), list(self._prev_layers), lambda *ts: torch.cat(ts, dim=ax),
                              self._output_shape)

    @property
    def output(self):
        return _logical(self.act.data, self._output_shape)


class Conv3DLSTMLayer(Conv3DLayer):
    """lib/layers.py:508-575.  The reference class is a stub: it computes an unused gate filter shape and then
    exactly Conv3DLayer's output (zero pad + conv3d2d.conv3d(W) + b); the conv-LSTM *network* of BASELINE.json
    config 4 is models.ResidualLSTMNet."""

    def __init__(self, prev_layer, filter_shape, padding=None, params=None):
        super().__init__(prev_layer, filter_shape, padding, params)
        n_c, n_x = filter_shape[0], self._input_shape[2]
        self._gate_filter_shape = [4 * n_c, 1, n_x + n_c, 1, 1]


class BlockDiagonalLayer(Layer):
    """lib/layers.py:164-203: out[n, d.., o] = sum_k x[n, d.., k] * W[d.., k, o] (+ b[d.., o]) -- one small GEMM
    per position d, each a C-ABI contraction call (forward, weight gradient, data gradient)."""

    def __init__(self, prev_layer, n_out, params=None, bias=True):
        super().__init__(prev_layer)
        self._bias = bias
        self._output_shape = list(self._input_shape)
        self._output_shape[-1] = n_out
        self._output_shape = tuple(self._output_shape)
        if params is None:
            self._W_shape = tuple(list(self._input_shape[1:]) + [n_out])
            self.W = Weight(self._W_shape, is_bias=False)
            if bias:
                self.b = Weight(self._output_shape[1:], is_bias=True, mean=0.1, filler='constant')
        else:
            self.W = params[0]
            if bias:
                self.b = params[1]
        self.W.bind(('blockdiag',))
        self.params = [self.W]
        if bias:
            self.params.append(self.b)

    def set_output(self):
        ctx = get_context()
        x = self._prev_layer.act
        ishape = tuple(int(v) for v in self._input_shape)
        N, K, O = ishape[0], ishape[-1], int(self._output_shape[-1])
        P = int(np.prod(ishape[1:-1]))
        W, b = self.W.param(ctx), (self.b.param(ctx) if self._bias else None)
        xt = _logical(x.data, ishape).reshape(N, P, K).transpose(0, 1).contiguous()          # [P,N,K]
        y = torch.empty(P, N, O, dtype=x.data.dtype, device=x.data.device)
        wop = W.operand().view(P, O * K)
        for p in range(P):
            ops.conv_fwd(xt[p].reshape(-1), wop[p], None if b is None else b.data[p * O:(p + 1) * O], None, None,
                         y[p].view(-1), (N, 1, 1, 1, K), O, (1, 1, 1), L.EPI_BIAS if b is not None else 0, ctx.algo)
        oshape = tuple(int(v) for v in self._output_shape)
        out = Act(_physical(y.transpose(0, 1).reshape(oshape), oshape), _geom_of(oshape), requires_grad=ctx.training)
        self._act = out
        if not ctx.training:
            return
        ctx.consume(x)

        def bwd():
            if out.grad is None:
                return
            dy = _logical(out.grad, oshape).reshape(N, P, O).transpose(0, 1).contiguous()   # [P,N,O]
            dx = torch.empty(P, N, K, dtype=x.data.dtype, device=x.data.device)
            wt = torch.empty(K * O, dtype=W.operand().dtype, device=x.data.device)
            for p in range(P):
                ops.conv_wgrad(xt[p].reshape(-1), dy[p].reshape(-1), W.grad[p * O * K:(p + 1) * O * K],
                               (N, 1, 1, 1, K), O, (1, 1, 1), ctx.wgrad_algo, W.grad_written)
                if b is not None:
                    ops.bias_grad(dy[p].reshape(-1), b.grad[p * O:(p + 1) * O], O, b.grad_written)
                if x.requires_grad:
                    ops.weight_transpose(W.data[p * O * K:(p + 1) * O * K], wt, O, 1, K)
                    ops.conv_fwd(dy[p].reshape(-1), wt, None, None, None, dx[p].view(-1), (N, 1, 1, 1, O), K,
                                 (1, 1, 1), 0, ctx.algo)
            W.grad_written = True
            if b is not None:
                b.grad_written = True
            if x.requires_grad:
                ctx.accumulate(x, _physical(dx.transpose(0, 1).reshape(ishape), ishape))
            out.grad = None
        ctx.tape.append(bwd)

    @property
    def output(self):
        return _logical(self.act.data, self._output_shape)

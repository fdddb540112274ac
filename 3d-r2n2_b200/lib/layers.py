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
        return self._owner.np_values

    def set_value(self, v):
        v = np.asarray(v, dtype=np.float32)
        if tuple(v.shape) != tuple(np.shape(self._owner.np_values)):
            raise ValueError('shape mismatch: %s vs %s' % (v.shape, np.shape(self._owner.np_values)))
        self._owner.np_values = v
        self._owner._param = None      # re-pack lazily


class Weight(object):
    """lib/layers.py:18-79: same signature, same fillers, same fan-in/out rules."""

    def __init__(self, w_shape, is_bias, mean=0, std=0.01, filler='msra', fan_in=None,
                 fan_out=None, name=None):
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
        global trainable_params
        trainable_params.append(self)

    # ---- packed device operand (built lazily, rebuilt after set_value) -----------------
    def bind(self, kind):
        if self._kind is None:
            self._kind = kind

    def param(self, ctx):
        if self._param is not None and self._param_dtype == ctx.dtype:
            return self._param
        t = torch.as_tensor(self.np_values).to(DEVICE)
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
        else:
            raise ValueError(kind)
        p = p.contiguous().view(-1)
        mirror = None
        if ctx.dtype != torch.float32 and not self.is_bias:
            mirror = torch.empty_like(p, dtype=ctx.dtype)
            ops.cast(p, mirror)
        self._param = Param(p, None, mirror, cout, taps, cin, None)
        self._param_dtype = ctx.dtype
        return self._param


def _cin_pad(ctx, cin):
    if cin % 4 == 0 and (ctx.algo == L.ALGO_SIMT or cin % 16 == 0):
        return cin
    return (cin + 3) // 4 * 4 if ctx.algo == L.ALGO_SIMT else (cin + 15) // 16 * 16


def _as_act(t, shape, ctx):
    """Reference-layout tensor / ndarray (or an Act) -> channels-last Act."""
    if isinstance(t, Act):
        return t
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
    raise ValueError('unsupported input rank %d' % t.dim())


def _ref_view(act, shape):
    """channels-last Act -> tensor with the reference's logical shape."""
    N, D, H, W, C = act.geom
    d = act.data.view(N, D, H, W, C)
    if len(shape) == 4:
        return d[:, 0].permute(0, 3, 1, 2)[:, :shape[1]]
    if len(shape) == 5:
        return d.permute(0, 1, 4, 2, 3)
    if len(shape) == 2:
        return d.reshape(N, -1)
    raise ValueError(shape)


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
        xin = Act(x.data, (N, 1, 1, 1, x.data.numel() // N))
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
        a, b = self._prev_layer.act, self._mult_layer.act
        y = torch.empty_like(a.data)
        ops.eltwise(L.ELT_MUL, a.data, b.data, y)
        self._act = Act(y, a.geom)


class FlattenLayer(Layer):
    """lib/layers.py:228-236: flatten(2) is (C,H,W)-major in the reference; here the buffer stays
    channels-last and the consumer's weight rows are permuted at pack time (`_chw`)."""

    def __init__(self, prev_layer):
        super().__init__(prev_layer)
        self._output_shape = [self._input_shape[0], int(np.prod(self._input_shape[1:]))]
        self._chw = tuple(int(s) for s in self._input_shape[1:]) if len(self._input_shape) == 4 else None

    def set_output(self):
        x = self._prev_layer.act
        self._act = Act(x.data, (x.geom[0], 1, 1, 1, x.data.numel() // x.geom[0]),
                        lrelu_out=x.lrelu_out)

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
        fin = Act(f.data, (B, 1, 1, 1, f.data.numel() // B))
        xp = ctx.conv(fin, self.Wx.param(ctx), None, (1, 1, 1))
        xp = Act(xp.data, (h.geom[0], h.geom[1], h.geom[2], h.geom[3], self._output_shape[2]))
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
        if input.data.dtype != torch.float32:
            input = Act(input.data.float(), input.geom)
        self.input = input
        B, D, H, W, C = input.geom
        if C != 2 or not (D == H == W):
            raise ValueError('SoftmaxWithLoss3D is accelerated for (B,n,2,n,n) inputs only')
        self._B, self._nv = B, D

    def _run(self, y, want_pred):
        B, nv = self._B, self._nv
        dev = self.input.data.device
        pred = torch.empty(B, nv, 2, nv, nv, dtype=torch.float32, device=dev) if want_pred else None
        loss = None
        yt = None
        if y is not None:
            yt = torch.as_tensor(y).to(dev).float().contiguous()
            loss = torch.zeros(1, dtype=torch.float32, device=dev)
        ops.softmax_ce3d(self.input.data, yt, pred, loss, None, B, nv, 1.0)
        return pred, (None if loss is None else loss[0] / float(B * nv ** 3))

    def prediction(self):
        return self._run(None, True)[0]

    def error(self, y, threshold=0.5):
        p = self.prediction()
        yt = torch.as_tensor(y).to(p.device).float()
        return torch.mean(((p >= threshold).float() == yt).float())

    def loss(self, y):
        return self._run(y, False)[1]


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
    _op = None

    def __init__(self, prev_layer):
        super().__init__(prev_layer)
        self._output_shape = self._input_shape

    def set_output(self):
        a = self._prev_layer.act
        y = torch.empty_like(a.data)
        ops.eltwise(self._op, a.data, None, y)
        self._act = Act(y, a.geom)


class SigmoidLayer(_Elt):
    """lib/layers.py:649-656"""
    _op = L.ELT_SIGMOID


class TanhLayer(_Elt):
    """lib/layers.py:659-665 (the reference forgets to set _output_shape there)"""
    _op = L.ELT_TANH


class ComplementLayer(_Elt):
    """lib/layers.py:668-676: 1 - input"""
    _op = L.ELT_COMPLEMENT

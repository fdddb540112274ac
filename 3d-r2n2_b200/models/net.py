"""Mirror of models/net.py: `Net` owns the parameters, the compiled (here: fused-engine)
forward/backward and save/load in the reference's weights.npy format."""
import datetime as dt

import numpy as np
import torch

from ..lib.config import cfg
from ..lib import theano_compat as theano
from ..lib import layers as ly
from ..engine import Engine
from ..ops import Ctx
from .. import _lib as L

tensor = theano.tensor
tensor5 = tensor.TensorType(theano.config.floatX, (False,) * 5)     # models/net.py:9


class _SharedView:
    """theano.shared stand-in backed by a packed view of the engine's flat store."""

    def __init__(self, net, ref_weight):
        self._net, self._w = net, ref_weight

    def get_value(self):
        return self._w.get_value()

    def set_value(self, v):
        self._w.set_value(v)
        self._net.engine.params_changed()


class NetWeight:
    """What `net.params` holds: the reference's Weight surface (lib/layers.py:66-76)."""

    def __init__(self, net, ref_weight):
        self.val = _SharedView(net, ref_weight)
        self.shape = ref_weight.ref_shape
        self.is_bias = ref_weight.is_bias
        self.name = ref_weight.name


class Net(object):
    """Two kinds of subclass:
      * the shipped networks set NET_NAME and run on the fused engine (engine.py);
      * a subclass that overrides `network_definition` the way the reference's model files do
        (models/net.py:43-51: builds layers from self.x / self.y, sets self.loss / error / params / output /
        activations) runs DEFINE-BY-RUN: the body is executed once at construction (creating the Weights, on a
        zero batch) and again on every call with the real batch fed through self.x / self.y, Weight
        constructions replaying the tensors of the first run (lib.layers.replay_weights); `tensor.grad` is the
        reverse pass over the tape the layers recorded (models/net.py:56-58)."""
    NET_NAME = None   # engine network name, set by subclasses

    def __init__(self, random_seed=dt.datetime.now().microsecond, compute_grad=True):
        self.rng = np.random.RandomState(random_seed)
        self.batch_size = cfg.CONST.BATCH_SIZE
        self.img_w = cfg.CONST.IMG_W
        self.img_h = cfg.CONST.IMG_H
        self.n_vox = cfg.CONST.N_VOX
        self.compute_grad = compute_grad
        self.engine = None
        self._feed = {'x': None, 'y': None}
        self._sym = {'x': None, 'y': None}
        # (self.batch_size, 3, self.img_h, self.img_w); a multi-view network re-declares x as tensor5 and
        # clears is_x_tensor4 in its network_definition (models/net.py:23-29)
        self.x = tensor.tensor4()
        self.is_x_tensor4 = True
        self.y = tensor5()
        self.activations = []
        self.loss = []
        self.output = []
        self.error = []
        self.params = []
        self.grads = []
        self.setup()

    # ---- self.x / self.y: symbolic declarations in the reference, the fed batch here ----------------------
    def _get_input(self, k):
        if self._feed[k] is None:
            sym = self._sym[k]
            if k == 'y' or sym is None:
                shape = (self.batch_size, self.n_vox, 2, self.n_vox, self.n_vox)
            elif sym.ndim == 5:
                shape = (1, self.batch_size, 3, self.img_h, self.img_w)
            else:
                shape = (self.batch_size, 3, self.img_h, self.img_w)
            self._feed[k] = torch.zeros(shape, dtype=torch.float32, device=ly.DEVICE)   # construction-time batch
        return self._feed[k]

    def _set_input(self, k, v):
        if isinstance(v, theano.Symbol):
            self._sym[k] = v             # a (re-)declaration: the fed batch stays
        else:
            self._feed[k] = v

    x = property(lambda self: self._get_input('x'), lambda self, v: self._set_input('x', v))
    y = property(lambda self: self._get_input('y'), lambda self, v: self._set_input('y', v))

    def setup(self):
        self._generic = type(self).network_definition is not Net.network_definition
        if self._generic:
            self._run_definition(None, None, False, first=True)
        else:
            self.network_definition()
        self.post_processing()

    def _run_definition(self, x, y, training, first=False):
        compute = cfg.CONST.COMPUTE
        if compute not in ('fp32', 'bf16'):
            raise NotImplementedError("user-defined networks run with CONST.COMPUTE 'fp32' or 'bf16'")
        dtype, algo = {'fp32': (torch.float32, L.ALGO_SIMT), 'bf16': (torch.bfloat16, L.ALGO_UMMA)}[compute]
        ctx = Ctx(dtype, algo, training=training)
        ly.set_context(ctx)
        self._feed = {'x': x, 'y': y}
        if first:
            n0 = len(ly.trainable_params)
            self.network_definition()
            self._weights = list(ly.trainable_params[n0:])
        else:
            if training:
                for w in self._weights:
                    if w._param is not None:
                        w._param.grad_written = False
            with ly.replay_weights(self._weights):
                self.network_definition()
        return ctx

    def network_definition(self):
        """Subclasses with a NET_NAME get the fused engine; anything else must override."""
        if self.NET_NAME is None:
            raise NotImplementedError("Virtual Function")
        self.is_x_tensor4 = False
        if self.img_w != self.img_h:
            raise NotImplementedError('square views only')
        self.engine = Engine(self.NET_NAME, self.batch_size, compute=cfg.CONST.COMPUTE,
                             img=self.img_w, n_vox=self.n_vox)
        self.params = [NetWeight(self, w) for w in self.engine.weights]
        self.init_weights()

    def init_weights(self):
        """lib/layers.py:33-64 initialisation (MSRA normal / constant 0.1), drawn from self.rng in
        parameter-creation order."""
        arrays = []
        for w in self.engine.weights:
            shape = w.ref_shape
            if w.is_bias:
                arrays.append(np.full(shape, 0.1, np.float32))
                continue
            fan_in = np.prod(shape[1:])
            fan_out = np.prod(shape) / (shape[2] if len(shape) == 5 else shape[1])
            n = (fan_in + fan_out) / 2.
            arrays.append(np.asarray(self.rng.normal(0, np.sqrt(2. / n), shape), dtype=np.float32))
        self.engine.set_params(arrays)

    def add_layer(self, layer):
        raise NotImplementedError("TODO: add a layer")

    def post_processing(self):
        # reference: self.grads = tensor.grad(self.loss, params) (symbolic).  Here gradients are
        # produced per call by forward_backward(); `grads` exposes the last ones.
        pass

    # ---- execution -----------------------------------------------------------------------
    def _to_dev(self, a):
        if a is None:
            return None
        t = torch.as_tensor(a)
        if t.dtype != torch.float32:
            t = t.float()
        if self.engine is None:
            return t.to(ly.DEVICE)
        return t.cuda(non_blocking=True)

    def _stage_inputs(self, x, y):
        """Host -> device copies of a training batch on a side stream, one event per view: the engine runs
        the first layers view by view behind the copies (44 MB at PCIe speed is ~1.8 ms of a ~12 ms step),
        and y is only needed by the loss at the very end of the forward pass."""
        xt = torch.as_tensor(x)
        yt = torch.as_tensor(y)
        if xt.dtype != torch.float32:
            xt = xt.float()
        if yt.dtype != torch.float32:
            yt = yt.float()
        if getattr(self, '_copy_stream', None) is None:
            self._copy_stream = torch.cuda.Stream()
        cur = torch.cuda.current_stream()
        xd = torch.empty(xt.shape, dtype=torch.float32, device='cuda')
        yd = torch.empty(yt.shape, dtype=torch.float32, device='cuda')
        side = self._copy_stream
        side.wait_stream(cur)                    # the buffers may recycle memory still in use on `cur`
        events = []
        with torch.cuda.stream(side):
            for t in range(xt.shape[0]):
                xd[t].copy_(xt[t], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(side)
                events.append(ev)
            yd.copy_(yt, non_blocking=True)
            yev = torch.cuda.Event()
            yev.record(side)
        # allocated on `cur`, written on `side`: tell the caching allocator so that a later free cannot hand
        # the blocks to `cur` while a copy is still in flight
        xd.record_stream(side)
        yd.record_stream(side)
        return xd, yd, events, yev

    def _generic_out(self, training):
        gates = None
        if self.activations:
            a = self.activations[0]
            gates = a.stack() if hasattr(a, 'stack') else a
        return dict(pred=None if training else self.output, loss=self.loss, update_gates=gates)

    def forward(self, x, y=None):
        """(prediction, loss, activations) as CUDA tensors; the compiled test function of
        lib/solver.py:215-218."""
        if self._generic:
            self._run_definition(self._to_dev(x), self._to_dev(y), False)
            out = self._generic_out(False)
            if y is None:
                out['loss'] = None
            return out
        xd, yd = self._to_dev(x), self._to_dev(y)
        if cfg.CONST.INFER_GRAPH and xd.is_cuda:
            out = self.engine.forward_graph(xd.contiguous(), None if yd is None else yd.contiguous())
        else:
            out = self.engine.forward(xd, yd, training=False)
        self.output, self.loss, self.activations = out['pred'], out['loss'], [out['update_gates']]
        return out

    def forward_backward(self, x, y):
        """loss + tensor.grad(loss, params) (models/net.py:58); gradients stay on the device in
        the engine's flat buffer; `self.grads` materialises them in reference layout lazily."""
        if not self.compute_grad:
            raise RuntimeError('net was built with compute_grad=False')
        if self._generic:
            ctx = self._run_definition(self._to_dev(x), self._to_dev(y), True)
            ctx.backward()
            for w in self._weights:          # a parameter the loss does not depend on: zero gradient
                if w._param is not None and not w._param.grad_written:
                    w._param.grad.zero_()
            return self._generic_out(True)
        xe = ye = None
        if cfg.CONST.OVERLAP_H2D and self.engine.store.p.is_cuda and not (torch.is_tensor(x) and x.is_cuda):
            x, y, xe, ye = self._stage_inputs(x, y)
        out = self.engine.forward(self._to_dev(x), self._to_dev(y), training=True, want_pred=False,
                                  x_events=xe, y_event=ye)
        self.engine.backward()
        self.loss = out['loss']
        return out

    def get_grads(self):
        if self._generic:
            return [w.grad_value() for w in self._weights]
        return self.engine.get_grads()

    # ---- checkpoint format of the reference ------------------------------------------------
    def save(self, filename):
        """models/net.py:60-67: np.save of the list of parameter arrays in creation order."""
        params_cpu = [param.val.get_value() for param in self.params]
        arr = np.empty(len(params_cpu), dtype=object)
        for i, p in enumerate(params_cpu):
            arr[i] = p
        np.save(filename, arr, allow_pickle=True)
        print('saving network parameters to ' + filename)

    def load(self, filename, ignore_param=True):
        """models/net.py:69-86 (short files are tolerated when ignore_param is True)."""
        print('loading network parameters from ' + filename)
        params_cpu_file = np.load(filename, allow_pickle=True)
        if filename.endswith('npz'):
            params_cpu = params_cpu_file[list(params_cpu_file.keys())[0]]
        else:
            params_cpu = params_cpu_file
        succ_ind = 0
        for param_idx, param in enumerate(self.params):
            try:
                if self.engine is not None:
                    param.val._w.set_value(params_cpu[succ_ind])     # one operand refresh after the loop
                else:
                    param.val.set_value(params_cpu[succ_ind])
                succ_ind += 1
            except IndexError:
                if ignore_param:
                    print('Ignore mismatch')
                else:
                    raise
        if self.engine is not None:
            self.engine.params_changed()

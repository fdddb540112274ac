"""Mirror of models/net.py: `Net` owns the parameters, the compiled (here: fused-engine)
forward/backward and save/load in the reference's weights.npy format."""
import datetime as dt

import numpy as np
import torch

from ..lib.config import cfg
from ..engine import Engine
from ..ops import Ctx
from .. import _lib as L


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
    NET_NAME = None   # engine network name, set by subclasses

    def __init__(self, random_seed=dt.datetime.now().microsecond, compute_grad=True):
        self.rng = np.random.RandomState(random_seed)
        self.batch_size = cfg.CONST.BATCH_SIZE
        self.img_w = cfg.CONST.IMG_W
        self.img_h = cfg.CONST.IMG_H
        self.n_vox = cfg.CONST.N_VOX
        self.compute_grad = compute_grad
        self.is_x_tensor4 = True
        self.x = None            # fed per call (reference: symbolic tensor5)
        self.y = None
        self.activations = []
        self.loss = []
        self.output = []
        self.error = []
        self.params = []
        self.grads = []
        self.setup()

    def setup(self):
        self.network_definition()
        self.post_processing()

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

    def forward(self, x, y=None):
        """(prediction, loss, activations) as CUDA tensors; the compiled test function of
        lib/solver.py:215-218."""
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
        xe = ye = None
        if cfg.CONST.OVERLAP_H2D and self.engine.store.p.is_cuda and not (torch.is_tensor(x) and x.is_cuda):
            x, y, xe, ye = self._stage_inputs(x, y)
        out = self.engine.forward(self._to_dev(x), self._to_dev(y), training=True, want_pred=False,
                                  x_events=xe, y_event=ye)
        self.engine.backward()
        self.loss = out['loss']
        return out

    def get_grads(self):
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
                param.val._w.set_value(params_cpu[succ_ind])
                succ_ind += 1
            except IndexError:
                if ignore_param:
                    print('Ignore mismatch')
                else:
                    raise
        self.engine.params_changed()

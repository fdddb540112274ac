"""Generate golden fixtures by importing the REAL reference sources from /root/reference.

Run in the build container only (the reference does not travel to the GPU box):
    python tests/golden/make_reference_fixtures.py

What can be pinned against real reference code, given that Theano is absent:
  * lib/voxel.py:evaluate_voxel_prediction  -> voxel_eval.npz   (pure numpy, runs as is)
  * lib/layers.py constructors (shape inference + parameter creation, pure Python/numpy)
    and models/{gru_net,res_gru_net}.py:network_definition (parameter creation ORDER,
    shapes, is_bias flags, init statistics, every layer's output_shape in construction
    order)                                   -> ref_structure.json
The symbolic Theano ops themselves (`tensor.nnet.conv2d`, `conv3d2d.conv3d`, `pool_2d`,
`scan`, `grad`) cannot run: they are replaced by inert mocks purely so that the reference
modules import and their constructors execute.  No arithmetic of the hot path is produced
by this script; that is why DESIGN.md says "parity unpinned" for the numerics.

Environment shims (not changes to reference logic): `collections.Iterable` (removed in
py3.10, used at lib/layers.py:33), `np.cast` (removed in numpy 2, used at
lib/layers.py:59), `easydict` (not installed; lib/config.py:1).
"""
import collections
import collections.abc
import json
import os
import sys
import types
from unittest import mock

import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))


def install_shims():
    collections.Iterable = collections.abc.Iterable
    if not hasattr(np, 'cast'):
        class _Cast(dict):
            def __getitem__(self, k):
                return lambda a: np.asarray(a, dtype=k)
        np.cast = _Cast()

    class EasyDict(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v
    ed = types.ModuleType('easydict')
    ed.EasyDict = EasyDict
    sys.modules['easydict'] = ed

    # inert theano: only `config.floatX` and `shared` carry meaning
    th = mock.MagicMock(name='theano')
    th.config.floatX = 'float32'

    class Shared:
        def __init__(self, value=None, **kw):
            self._v = np.asarray(value)

        def get_value(self):
            return self._v

        def set_value(self, v):
            self._v = np.asarray(v)

        def dimshuffle(self, *a):
            return mock.MagicMock()
    th.shared = Shared

    def scan(fn, sequences=None, outputs_info=None, **kw):
        # run the step function once so the per-step layers get constructed/recorded
        fn(mock.MagicMock(), mock.MagicMock(), mock.MagicMock())
        return mock.MagicMock(), mock.MagicMock()
    th.scan = scan
    for name in ['theano', 'theano.tensor', 'theano.tensor.nnet', 'theano.tensor.signal',
                 'theano.tensor.nnet.conv', 'theano.tensor.nnet.conv3d2d',
                 'theano.tensor.signal.pool', 'theano.sandbox', 'theano.sandbox.cuda']:
        sys.modules[name] = th if name == 'theano' else mock.MagicMock(name=name)
    sys.modules['theano'].tensor = sys.modules['theano.tensor']


def main():
    install_shims()
    sys.path.insert(0, REF)
    import lib.voxel as ref_voxel
    import lib.layers as ref_layers
    from lib.config import cfg

    # ---- 1. evaluate_voxel_prediction golden -------------------------------------------
    rng = np.random.RandomState(7)
    preds = rng.rand(6, 32, 2, 32, 32).astype(np.float32)
    preds[:, :, 0] = 1.0 - preds[:, :, 1]
    occ = rng.rand(6, 32, 32, 32) < 0.15
    gt = np.zeros((6, 32, 2, 32, 32), np.float32)
    gt[:, :, 1] = occ
    gt[:, :, 0] = ~occ
    thresholds = [0.2, 0.4, 0.5, 0.9]
    counts = np.zeros((6, len(thresholds), 5), np.int64)
    for b in range(6):
        for ti, th in enumerate(thresholds):
            counts[b, ti] = ref_voxel.evaluate_voxel_prediction(preds[b], gt[b].astype(bool), th)
    np.savez_compressed(os.path.join(HERE, 'voxel_eval.npz'),
                        preds1=preds[:, :, 1].astype(np.float16),   # small fixture: fp16 grid
                        occ=np.packbits(occ), thresholds=np.array(thresholds), counts=counts)
    # recompute with the fp16-rounded predictions so the fixture is self-consistent
    p16 = preds[:, :, 1].astype(np.float16).astype(np.float32)
    for b in range(6):
        pr = np.stack([1 - p16[b], p16[b]], axis=1)
        for ti, th in enumerate(thresholds):
            counts[b, ti] = ref_voxel.evaluate_voxel_prediction(pr, gt[b].astype(bool), th)
    np.savez_compressed(os.path.join(HERE, 'voxel_eval.npz'),
                        preds1=p16.astype(np.float16), occ=np.packbits(occ),
                        thresholds=np.array(thresholds), counts=counts)

    # ---- 2. structure: record every layer construction ------------------------------------
    record = []

    def wrap(cls):
        orig = cls.__init__

        def init(self, *a, **kw):
            orig(self, *a, **kw)
            if type(self) is cls:
                try:
                    shp = [int(s) for s in self.output_shape]
                except Exception:
                    shp = None
                record.append([cls.__name__, shp])
        cls.__init__ = init

    layer_classes = ['InputLayer', 'TensorProductLayer', 'ConvLayer', 'PoolLayer', 'Unpool3DLayer',
                     'Conv3DLayer', 'FCConv3DLayer', 'LeakyReLU', 'SigmoidLayer', 'TanhLayer',
                     'ComplementLayer', 'AddLayer', 'EltwiseMultiplyLayer', 'FlattenLayer']
    for n in layer_classes:
        wrap(getattr(ref_layers, n))

    out = {'cfg': {'BATCH_SIZE': cfg.CONST.BATCH_SIZE, 'IMG_W': cfg.CONST.IMG_W,
                   'IMG_H': cfg.CONST.IMG_H, 'N_VOX': cfg.CONST.N_VOX, 'N_VIEWS': cfg.CONST.N_VIEWS,
                   'WEIGHT_DECAY': cfg.TRAIN.WEIGHT_DECAY, 'LR': cfg.TRAIN.DEFAULT_LEARNING_RATE,
                   'VOXEL_THRESH': cfg.TEST.VOXEL_THRESH, 'POLICY': cfg.TRAIN.POLICY},
           'nets': {}}
    import models.gru_net as ref_gru
    import models.res_gru_net as ref_res
    for name, mod in [('GRUNet', ref_gru), ('ResidualGRUNet', ref_res)]:
        del record[:]
        del ref_layers.trainable_params[:]      # quirk 13: global list never cleared
        np.random.seed(0)
        net = getattr(mod, name)(compute_grad=False)
        params = []
        for w in net.params:
            v = w.np_values
            params.append({'shape': [int(s) for s in np.shape(v)], 'is_bias': bool(w.is_bias),
                           'mean': float(np.mean(v)), 'std': float(np.std(v)),
                           'dtype': str(v.dtype)})
        out['nets'][name] = {'params': params, 'layers': list(record),
                             'n_params': int(sum(int(np.prod(p['shape'])) for p in params))}
        print(name, len(params), 'tensors', out['nets'][name]['n_params'], 'params',
              len(record), 'layers')
    with open(os.path.join(HERE, 'ref_structure.json'), 'w') as f:
        json.dump(out, f, indent=0)


if __name__ == '__main__':
    main()

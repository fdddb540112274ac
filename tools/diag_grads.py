"""Diagnostic: per-parameter gradient error of the fp32 CUDA path vs the fp64 / fp32 oracle."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import r2n2_b200  # noqa
from r2n2_b200.lib.config import cfg
from r2n2_b200.models import load_model
from oracle import ref_torch as rt
B, T = 2, 2
cfg.CONST.BATCH_SIZE = B
net = load_model('ResidualGRUNet')()
params = rt.init_params('ResidualGRUNet', 0)
net.engine.set_params(params)
x, y = rt.synthetic_batch(T, B)
net.forward_backward(x, y)
g = net.get_grads()
r64 = rt.RefNet('ResidualGRUNet', params, dtype=torch.float64); _, g64 = r64.loss_and_grads(x, y)
r32 = rt.RefNet('ResidualGRUNet', params, dtype=torch.float32); _, g32 = r32.loss_and_grads(x, y)
l2 = lambda a, b: np.linalg.norm((a - b).ravel()) / (np.linalg.norm(b.ravel()) + 1e-30)
for (name, _, _), a, b, c in zip(r64.spec, g, g64, g32):
    b = b.numpy(); c = c.numpy().astype(np.float64)
    print('%-18s gpu-vs-64 %.2e   o32-vs-64 %.2e   gpu-vs-o32 %.2e' % (name, l2(a, b), l2(c, b), l2(a, c)))

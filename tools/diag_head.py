#!/usr/bin/env python
"""Diagnostic: the bf16 training forward with HOST inputs (chunked head behind the H2D copies) against the same
forward with device-resident inputs, stage by stage, and run-to-run determinism of each."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import r2n2_b200  # noqa: F401
from r2n2_b200.lib.config import cfg
from r2n2_b200.models import load_model

B, T = int(os.environ.get('B', 36)), int(os.environ.get('T', 5))
cfg.CONST.BATCH_SIZE = B
cfg.CONST.COMPUTE = 'bf16'
net = load_model('ResidualGRUNet')(random_seed=0)
eng = net.engine
rs = np.random.RandomState(3)
x = rs.uniform(0, 1, (T, B, 3, 127, 127)).astype(np.float32)
occ = rs.rand(B, 32, 32, 32) < 0.1
y = np.zeros((B, 32, 2, 32, 32), np.float32)
y[:, :, 1] = occ
y[:, :, 0] = ~occ
xd, yd = torch.as_tensor(x).cuda(), torch.as_tensor(y).cuda()


def run(host):
    eng.debug = {}
    if host:
        out = net.forward_backward(x, y)
    else:
        out = net.forward_backward(xd, yd)
    torch.cuda.synchronize()
    d = {k: v.clone() for k, v in eng.debug.items()}
    d['loss'] = out['loss'].clone()
    d['g'] = eng.store.g.clone()
    return d


def cmp(a, b, tag):
    for k in a:
        u, v = a[k].float(), b[k].float()
        nd = int((u != v).sum())
        print('%-14s %-10s differing %9d / %-10d max|d| %.3e rel-l2 %.3e' % (
            tag, k, nd, u.numel(), float((u - v).abs().max()), float((u - v).norm() / (v.norm() + 1e-30))))


d1, d2 = run(False), run(False)
cmp(d1, d2, 'dev vs dev')
h1, h2 = run(True), run(True)
cmp(h1, h2, 'host vs host')
cmp(h1, d1, 'host vs dev')
# per-segment gradient differences host vs dev
st = eng.store
rows = []
for name, (off, n) in st.offsets.items():
    u, v = h1['g'][off:off + n], d1['g'][off:off + n]
    rows.append((float((u - v).norm() / (v.norm() + 1e-30)), name))
for r, name in sorted(rows, reverse=True)[:12]:
    print('grad %-14s rel-l2 %.3e' % (name, r))

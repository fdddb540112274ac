#!/usr/bin/env python
"""Where do the tcgen05 kernels wait?  DIAGNOSTIC tool (not part of the product path).

Builds / loads libr2n2_b200_ws.so (R2N2_WAITSTAT=1: csrc/umma.cuh counts the cycles every warp spends inside
mbar_wait, per barrier slot), runs one launch per layer shape and mode, and prints per warp ROLE the share of
the kernel's lifetime spent blocked on each barrier, averaged over the CTAs.

  R2N2_WAITSTAT=1 python tools/waitstat.py [--only conv1b,conv9b] [--what fwd,dgrad,wgrad] [--out file.txt]

Barrier slot order (8-byte slots from the kernel's first mbarrier):
  conv_fwd3_umma_kernel  : pfull[ring] pempty[ring] wfull[ws] wempty[ws] tfull[2] tempty[2] efull
  conv_fwd_umma_persistent: full[stages] empty[stages] tfull[2] tempty[2] efull...   (see the kernel)
  wgrad kernels          : full[stages] empty[stages] tmem_full
Warp roles: fwd3: 0 planes-TMA, 1 MMA, 2-9 epilogue, 10 weights-TMA; fwd: 0 TMA, 1 MMA, 2-9 epilogue;
wgrad: 0 TMA, 1 MMA, 2-5 epilogue.
"""
import argparse
import ctypes as C
import os
import sys

os.environ['R2N2_WAITSTAT'] = '1'
os.environ.setdefault('R2N2_UMMA_DBG', '1')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import numpy as np
import torch
import importlib.util
spec = importlib.util.spec_from_file_location('r2n2_build', os.path.join(ROOT, '3d-r2n2_b200', 'csrc', 'build.py'))
bld = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bld)
bld.build()
import r2n2_b200  # noqa: F401,E402
from r2n2_b200 import ops, _lib as L  # noqa: E402
from bench_conv import SHAPES  # noqa: E402

NW = 512 * 16 * 40


def read_stats():
    lib = L.load()
    buf = np.zeros(NW, np.int64)
    lib.r2n2_waitstat_read.argtypes = [C.c_void_p, C.c_longlong]
    rc = lib.r2n2_waitstat_read(buf.ctypes.data, NW)
    assert rc == 0, rc
    return buf.reshape(512, 16, 40)


def report(tag, st, ms, out):
    live = st[:, :, 38] > 0
    nb = int(live.any(1).sum())
    lines = ['== %s  %.4f ms  (%d CTAs)' % (tag, ms, nb)]
    for w in range(16):
        m = live[:, w]
        if not m.any():
            continue
        tot = st[m, w, 38].mean()
        bins = st[m, w, :38].mean(0)
        parts = ['b%d %.1f%%' % (i, 100 * bins[i] / tot) for i in range(38) if bins[i] > 0.005 * tot]
        lines.append('  warp %2d: life %8.0f cyc  blocked %5.1f%%  (%5.0f blocking waits)  %s' % (
            w, tot, 100 * bins.sum() / tot, st[m, w, 39].mean(), '  '.join(parts)))
    txt = '\n'.join(lines)
    print(txt, flush=True)
    if out:
        out.write(txt + '\n')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default=None)
    ap.add_argument('--what', default='fwd,dgrad,wgrad')
    ap.add_argument('--out', default=None)
    a = ap.parse_args()
    out = open(a.out, 'w') if a.out else None
    dev, bf = 'cuda', torch.bfloat16
    for name, (N, D, H, W), cin, cout, k in SHAPES:
        if a.only and name not in a.only.split(','):
            continue
        P, taps = N * D * H * W, k[0] * k[1] * k[2]
        x = torch.randn(P * cin, device=dev).to(bf)
        w = (torch.randn(cout * taps * cin, device=dev) * 0.05).to(bf)
        bias = torch.zeros(cout, device=dev)
        y = torch.empty(P * cout, device=dev, dtype=bf)
        dy = torch.randn(P * cout, device=dev).to(bf)
        dx = torch.empty(P * cin, device=dev, dtype=bf)
        dw = torch.zeros(cout * taps * cin, device=dev)
        colsum = torch.zeros(cin, device=dev)
        for what in a.what.split(','):
            if what == 'fwd':
                fn = lambda: ops.conv_fwd(x, w, bias, None, None, y, (N, D, H, W, cin), cout, k,
                                          L.EPI_BIAS | L.EPI_LRELU, L.ALGO_UMMA)
            elif what == 'dgrad':        # data gradient of the layer: dy [P,cout] -> dx [P,cin], LeakyReLU' mask + bias sums
                if name.endswith('_dgrad') or k == (1, 7, 1):
                    continue
                fn = lambda: ops.conv_fwd(dy, w, None, x, None, dx, (N, D, H, W, cout), cin, k, L.EPI_MASK, L.ALGO_UMMA,
                                          colsum)
            else:
                fn = lambda: ops.conv_wgrad(x, dy, dw, (N, D, H, W, cin), cout, k, L.ALGO_UMMA, False)
            fn(); fn()
            torch.cuda.synchronize()
            read_stats()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            report('%s %s  M=%d N=%d K=%d' % (name, what, P, cout if what != 'dgrad' else cin,
                                              taps * (cin if what != 'dgrad' else cout)), read_stats(), e0.elapsed_time(e1), out)
    assert ops.umma_error_flag() == 0


if __name__ == '__main__':
    main()

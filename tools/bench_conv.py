#!/usr/bin/env python
"""Micro-benchmark of the implicit-GEMM kernels on the layer shapes of ResidualGRUNet
(B=36, T=5).  CUDA-event timed, dense TFLOP/s (2*M*N*K / t) per shape.

  python tools/bench_conv.py [--only NAME] [--iters 5] [--what fwd,wgrad] [--json out.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import r2n2_b200  # noqa: F401,E402
from r2n2_b200 import ops, _lib as L  # noqa: E402

TB, B = 180, 36
SHAPES = [
    # name, (N,D,H,W), Cin, Cout, k
    ('conv1a', (TB, 1, 127, 127), 32, 96, (1, 7, 1)),      # row-packed first layer
    ('conv1b', (TB, 1, 127, 127), 96, 96, (1, 3, 3)),
    ('conv2a', (TB, 1, 64, 64), 96, 128, (1, 3, 3)),
    ('conv2b', (TB, 1, 64, 64), 128, 128, (1, 3, 3)),
    ('conv2a_dgrad', (TB, 1, 64, 64), 128, 96, (1, 3, 3)),
    ('conv2c', (TB, 1, 64, 64), 96, 128, (1, 1, 1)),
    ('conv3a', (TB, 1, 33, 33), 128, 256, (1, 3, 3)),
    ('conv3b', (TB, 1, 33, 33), 256, 256, (1, 3, 3)),
    ('conv4a', (TB, 1, 17, 17), 256, 256, (1, 3, 3)),
    ('conv5a', (TB, 1, 9, 9), 256, 256, (1, 3, 3)),
    ('conv6a', (TB, 1, 5, 5), 256, 256, (1, 3, 3)),
    ('fc7', (TB, 1, 1, 1), 2304, 1024, (1, 1, 1)),
    ('xproj_ur', (TB, 1, 1, 1), 1024, 16384, (1, 1, 1)),
    ('xproj_ur_dgrad', (TB, 1, 1, 1), 16384, 1024, (1, 1, 1)),
    ('gru_wh_ur', (B, 4, 4, 4), 128, 256, (3, 3, 3)),
    ('conv7a', (B, 8, 8, 8), 128, 128, (3, 3, 3)),
    ('conv8a', (B, 16, 16, 16), 128, 128, (3, 3, 3)),
    ('conv9a', (B, 32, 32, 32), 128, 64, (3, 3, 3)),
    ('conv9a_dgrad', (B, 32, 32, 32), 64, 128, (3, 3, 3)),
    ('conv9b', (B, 32, 32, 32), 64, 64, (3, 3, 3)),
    ('conv9c', (B, 32, 32, 32), 128, 64, (1, 1, 1)),
    ('conv10a', (B, 32, 32, 32), 64, 32, (3, 3, 3)),
    ('conv10a_dgrad', (B, 32, 32, 32), 32, 64, (3, 3, 3)),
    ('conv10b', (B, 32, 32, 32), 32, 32, (3, 3, 3)),
    ('conv11', (B, 32, 32, 32), 32, 16, (3, 3, 3)),
    ('conv11_dgrad', (B, 32, 32, 32), 16, 32, (3, 3, 3)),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default=None)
    ap.add_argument('--iters', type=int, default=5)
    ap.add_argument('--what', default='fwd,wgrad')
    ap.add_argument('--json', default=None)
    ap.add_argument('--graph', action='store_true', help='time the launches replayed from a CUDA graph (no host launch overhead)')
    args = ap.parse_args()
    dev = 'cuda'
    bf = torch.bfloat16
    out = {}
    for name, (N, D, H, W), cin, cout, k in SHAPES:
        if args.only and name not in args.only.split(','):
            continue
        P = N * D * H * W
        taps = k[0] * k[1] * k[2]
        x = torch.randn(P * cin, device=dev).to(bf)
        w = (torch.randn(cout * taps * cin, device=dev) * 0.05).to(bf)
        bias = torch.zeros(cout, device=dev)
        y = torch.empty(P * cout, device=dev, dtype=bf)
        dy = torch.randn(P * cout, device=dev).to(bf)
        dw = torch.zeros(cout * taps * cin, device=dev)
        flops = 2.0 * P * cin * cout * taps
        res = {}
        for what in args.what.split(','):
            if what == 'fwd':
                fn = lambda: ops.conv_fwd(x, w, bias, None, None, y, (N, D, H, W, cin), cout, k,
                                          L.EPI_BIAS | L.EPI_LRELU, L.ALGO_UMMA)
            elif what == 'dgrad':
                # data-gradient form: no bias, LeakyReLU' mask of the producing layer's output in the epilogue
                fn = lambda: ops.conv_fwd(x, w, None, dy, None, y, (N, D, H, W, cin), cout, k, L.EPI_MASK, L.ALGO_UMMA)
            else:
                fn = lambda: ops.conv_wgrad(x, dy, dw, (N, D, H, W, cin), cout, k, L.ALGO_UMMA, False)
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if args.graph:
                # small layers: a python launch (~20 us) is longer than the kernel; replay the launches from a CUDA graph
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    fn()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(args.iters):
                        fn()
                g.replay()
                torch.cuda.synchronize()
                e0.record()
                g.replay()
                e1.record()
            else:
                e0.record()
                for _ in range(args.iters):
                    fn()
                e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            res[what] = dict(ms=ms, tflops=flops / ms / 1e9)
        out[name] = res
        print('%-16s M=%-8d N=%-6d K=%-6d  ' % (name, P, cout, taps * cin) +
              '  '.join('%s %7.3f ms %7.1f TF/s' % (kk, v['ms'], v['tflops']) for kk, v in res.items()),
              flush=True)
    assert ops.umma_error_flag() == 0
    if args.json:
        json.dump(out, open(args.json, 'w'), indent=1)


if __name__ == '__main__':
    main()

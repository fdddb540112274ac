#!/usr/bin/env python
"""Micro-benchmark of the fused conv-GRU / conv-LSTM gate kernels and the other bandwidth-bound kernels at a batch
large enough to leave L2 ("large-batch micro-benchmark and the in-situ time").  In the network the
gate kernels move 3-5 MB per launch (B = 36) and are launch-latency bound; here B = 2304 (the x-projections of
64 steps of 36 samples) shows what the kernels do when they are bandwidth bound.  CUDA events, achieved GB/s =
algorithmic bytes (every operand read or written once) / time, against MEASURED_PEAKS.json.

  python tools/bench_gates.py [--batch 2304] [--iters 20] [--json out.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import r2n2_b200  # noqa: F401,E402
from r2n2_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=2304)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--json', default=None)
    a = ap.parse_args()
    peak = 6552.3
    try:
        peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        pass
    B, hn = a.batch, 64 * 128
    dev, bf = 'cuda', torch.bfloat16
    r = lambda n: torch.randn(n, device=dev).to(bf)
    conv_ur, xp_ur, h = r(B * 2 * hn), r(B * 2 * hn), r(B * hn)
    b_ur, b_c = torch.zeros(256, device=dev), torch.zeros(128, device=dev)
    u, rr, rh, c, hnew = (torch.empty(B * hn, device=dev, dtype=bf) for _ in range(5))
    conv_c, xp_c, dh = r(B * hn), r(B * hn), r(B * hn)
    da_ur, da_c, dhp, d_rh = torch.zeros(B * 2 * hn, device=dev, dtype=bf), torch.empty(B * hn, device=dev, dtype=bf), \
        torch.empty(B * hn, device=dev, dtype=bf), r(B * hn)
    fig, s_prev, s_new = torch.empty(B * 3 * hn, device=dev, dtype=bf), r(B * hn), torch.empty(B * hn, device=dev, dtype=bf)
    conv3, xp3, b3 = r(B * 3 * hn), r(B * 3 * hn), torch.zeros(384, device=dev)
    da3, dsp = torch.empty(B * 3 * hn, device=dev, dtype=bf), torch.empty(B * hn, device=dev, dtype=bf)
    el = 2
    cases = [
        ('gru_ur_fwd', lambda: ops.gru_ur_fwd(conv_ur, xp_ur, b_ur, h, u, rr, rh, B), (2 + 2 + 1 + 3) * hn * el),
        ('gru_out_fwd', lambda: ops.gru_out_fwd(conv_c, xp_c, b_c, u, h, c, hnew, B), (1 + 1 + 1 + 1 + 2) * hn * el),
        ('gru_out_bwd', lambda: ops.gru_out_bwd(dh, u, c, h, da_ur, da_c, dhp, B), (4 + 1 + 1 + 1) * hn * el),
        ('gru_r_bwd', lambda: ops.gru_r_bwd(d_rh, h, rr, da_ur, dhp, B), (3 + 1 + 2) * hn * el),
        ('lstm_fwd', lambda: ops.lstm_fwd(conv3, xp3, b3, s_prev, fig, s_new, hnew, B), (3 + 3 + 1 + 3 + 2) * hn * el),
        ('lstm_bwd', lambda: ops.lstm_bwd(dh, dhp, fig, s_prev, hnew, da3, dsp, B), (2 + 3 + 2 + 3 + 1) * hn * el),
    ]
    out = {}
    for name, fn, bytes_per_sample in cases:
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        gbs = bytes_per_sample * B / ms / 1e6
        out[name] = dict(ms=ms, gbs=gbs, frac=gbs / peak, mbytes=bytes_per_sample * B / 1e6)
        print('%-12s B=%d  %.1f MB  %.4f ms  %7.0f GB/s  %.2f of %.0f' % (name, B, bytes_per_sample * B / 1e6, ms, gbs, gbs / peak, peak),
              flush=True)
    if a.json:
        json.dump(out, open(a.json, 'w'), indent=1)


if __name__ == '__main__':
    main()

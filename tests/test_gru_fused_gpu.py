"""GPU parity of r2n2_conv_fwd_gru: the conv-GRU hidden convolutions with the gate math of
models/res_gru_net.py:130-151 (and of its gradient) in the tcgen05 epilogue, against the emulated C-ABI
semantics (float64 convolution + gate formulas on the same bf16 operands)."""
import numpy as np
import pytest
import torch

import r2n2_b200  # noqa: F401
from r2n2_b200 import ops, _lib as L
import emu

pytestmark = pytest.mark.gpu
DEV = 'cuda'
HC = 128


def _r(n, seed, scale=1.0):
    return torch.tensor(np.random.RandomState(seed).randn(n) * scale, dtype=torch.float32)


def _u(n, seed):
    return torch.tensor(np.random.RandomState(seed).rand(n), dtype=torch.float32)


def _run(mode, B, g=4):
    bf = torch.bfloat16
    P = B * g * g * g
    cin = 2 * HC if mode == L.GRU_OBWD else HC
    cout = 2 * HC if mode == L.GRU_UR else HC
    geom = (B, g, g, g, cin)
    x = _r(P * cin, 1).to(bf)
    w = _r(cout * 27 * cin, 2, (2.0 / (27 * cin)) ** 0.5).to(bf)
    bias = _r(cout, 3) if mode in (L.GRU_UR, L.GRU_BLEND) else None
    hn = P * HC
    if mode == L.GRU_UR:
        ins = [_r(P * 2 * HC, 4).to(bf), _r(hn, 5).to(bf)]
        outs = [torch.zeros(hn, dtype=bf) for _ in range(3)]
    elif mode == L.GRU_BLEND:
        ins = [_r(hn, 4).to(bf), _u(hn, 5).to(bf), _r(hn, 6).to(bf)]
        outs = [torch.zeros(hn, dtype=bf) for _ in range(2)]
    elif mode == L.GRU_RBWD:
        ins = [_r(hn, 4).to(bf), _u(hn, 5).to(bf)]
        outs = [_r(P * 2 * HC, 6).to(bf), _r(hn, 7).to(bf)]          # da_ur (u half must survive), dh_prev (in/out)
    else:
        ins = [_r(hn, 4).to(bf), _u(hn, 5).to(bf), _r(hn, 6, 0.5).clamp(-0.99, 0.99).to(bf), _r(hn, 7).to(bf)]
        outs = [torch.zeros(hn, dtype=bf), _r(P * 2 * HC, 8).to(bf), torch.zeros(hn, dtype=bf)]
    return geom, cout, x, w, bias, ins, outs


@pytest.mark.parametrize('mode', [L.GRU_UR, L.GRU_BLEND, L.GRU_RBWD, L.GRU_OBWD])
@pytest.mark.parametrize('B', [1, 3, 36, 40])
@pytest.mark.parametrize('no_h', [False, True])
def test_conv_fwd_gru(mode, B, no_h):
    if no_h and mode != L.GRU_OBWD:
        pytest.skip('only the out-gate gradient has an optional operand (h of the first view)')
    geom, cout, x, w, bias, ins, outs = _run(mode, B)
    if no_h:
        ins[3] = None
    want = [o.clone() for o in outs]
    emu.conv_fwd_gru(mode, x.float(), w.float(), bias, [None if t is None else t.float() for t in ins], want, geom,
                     cout, (3, 3, 3))
    got = [o.to(DEV) for o in outs]
    ops.conv_fwd_gru(mode, x.to(DEV), w.to(DEV), None if bias is None else bias.to(DEV),
                     [None if t is None else t.to(DEV) for t in ins], got, geom, cout, (3, 3, 3))
    torch.cuda.synchronize()
    assert ops.umma_error_flag() == 0, 'tcgen05 pipeline timed out'
    for i, (g_, w_) in enumerate(zip(got, want)):
        g_, w_ = g_.float().cpu(), w_.float()
        assert torch.isfinite(g_).all()
        err = (g_ - w_).abs().max().item()
        ref = w_.abs().max().item()
        # bf16 storage of the result (2^-8 relative) + fp32 accumulation order + __expf / tanhf
        assert err <= 1.2e-2 * max(1.0, ref), 'mode %d output %d: max err %.3e (ref %.3e)' % (mode, i, err, ref)


def test_conv_fwd_gru_rejects_fp32():
    geom, cout, x, w, bias, ins, outs = _run(L.GRU_BLEND, 1)
    with pytest.raises(RuntimeError):
        ops.conv_fwd_gru(L.GRU_BLEND, x.float().to(DEV), w.float().to(DEV), bias.to(DEV),
                         [t.float().to(DEV) for t in ins], [o.float().to(DEV) for o in outs], geom, cout, (3, 3, 3))

"""GPU parity of the fused Unpool3D + Conv3D path (lib/layers.py:375-385 + 441): R2N2_CONV_UP2 (only
the taps that meet a stored value), R2N2_CONV_DOWN2 (its data gradient: stride-2 TMA boxes) and
R2N2_WGRAD_UP2, against the emulated C-ABI semantics (dense convolution over the materialised
zero-stuffed tensor), bf16 operands."""
import numpy as np
import pytest
import torch

import r2n2_b200  # noqa: F401
from r2n2_b200 import ops, _lib as L
import emu

pytestmark = pytest.mark.gpu
DEV = 'cuda'
bf = torch.bfloat16


def _r(n, seed, scale=1.0):
    return torch.tensor(np.random.RandomState(seed).randn(n) * scale, dtype=torch.float32)


CASES = [
    # N, D, H, W (small grid), Cin, Cout, k
    (3, 4, 4, 4, 128, 128, (3, 3, 3)),     # conv7a
    (2, 8, 8, 8, 128, 128, (3, 3, 3)),     # conv8a
    (1, 16, 16, 16, 128, 64, (3, 3, 3)),   # conv9a
    (1, 16, 16, 16, 128, 64, (1, 1, 1)),   # conv9c: 7 of the 8 parity classes see no tap (bias only)
    (2, 3, 5, 6, 64, 32, (3, 3, 3)),       # ragged
    (1, 4, 4, 4, 32, 256, (3, 3, 3)),      # two N tiles of the weight gradient
]


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('flags,out_dtype', [(0, torch.float32), (L.EPI_BIAS | L.EPI_LRELU, bf),
                                             (L.EPI_BIAS | L.EPI_RESIDUAL, bf)])
def test_conv_up2_fwd(case, flags, out_dtype):
    N, D, H, W, Cin, Cout, k = case
    taps = k[0] * k[1] * k[2]
    Ps, Pb = N * D * H * W, 8 * N * D * H * W
    x = _r(Ps * Cin, 1).to(bf)
    w = _r(Cout * taps * Cin, 2, (2.0 / (taps * Cin)) ** 0.5).to(bf)
    b = _r(Cout, 3)
    res = _r(Pb * Cout, 4).to(out_dtype)
    want = torch.empty(Pb * Cout)
    emu.conv_fwd(x.float(), w.float(), b, None, res.float(), want, (N, D, H, W, Cin), Cout, k,
                 flags | L.CONV_UP2, 0)
    y = torch.full((Pb * Cout,), float('nan'), dtype=out_dtype, device=DEV)
    ops.conv_fwd(x.to(DEV), w.to(DEV), b.to(DEV), None, res.to(DEV), y, (N, D, H, W, Cin), Cout, k,
                 flags | L.CONV_UP2, L.ALGO_UMMA)
    torch.cuda.synchronize()
    assert ops.umma_error_flag() == 0
    got = y.float().cpu()
    assert torch.isfinite(got).all(), 'positions left unwritten'
    tol = 2e-3 if out_dtype == torch.float32 else 1.2e-2
    err = (got - want).abs().max().item()
    assert err <= tol * max(1.0, want.abs().max().item()), (case, flags, err)


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('flags', [0, L.EPI_RESIDUAL | L.EPI_MASK])
def test_conv_down2(case, flags):
    """Data gradient of an UP2 convolution: dy on the x2 grid (Cout channels) -> small grid (Cin)."""
    N, D, H, W, Cin, Cout, k = case
    taps = k[0] * k[1] * k[2]
    Ps, Pb = N * D * H * W, 8 * N * D * H * W
    dy = _r(Pb * Cout, 1).to(bf)
    wt = _r(Cin * taps * Cout, 2, (2.0 / (taps * Cout)) ** 0.5).to(bf)
    res = _r(Ps * Cin, 4).to(bf)
    mask = _r(Ps * Cin, 5).to(bf)
    want = torch.empty(Ps * Cin)
    emu.conv_fwd(dy.float(), wt.float(), None, mask.float(), res.float(), want, (N, D, H, W, Cout), Cin, k,
                 flags | L.CONV_DOWN2, 0)
    y = torch.full((Ps * Cin,), float('nan'), dtype=bf, device=DEV)
    ops.conv_fwd(dy.to(DEV), wt.to(DEV), None, mask.to(DEV), res.to(DEV), y, (N, D, H, W, Cout), Cin, k,
                 flags | L.CONV_DOWN2, L.ALGO_UMMA)
    torch.cuda.synchronize()
    assert ops.umma_error_flag() == 0
    got = y.float().cpu()
    assert torch.isfinite(got).all()
    err = (got - want).abs().max().item()
    assert err <= 1.2e-2 * max(1.0, want.abs().max().item()), (case, flags, err)


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('accumulate', [False, True])
def test_conv_wgrad_up2(case, accumulate):
    N, D, H, W, Cin, Cout, k = case
    taps = k[0] * k[1] * k[2]
    Ps, Pb = N * D * H * W, 8 * N * D * H * W
    x, dy = _r(Ps * Cin, 1).to(bf), _r(Pb * Cout, 2).to(bf)
    dw0 = _r(Cout * taps * Cin, 3)
    want = dw0.clone()
    emu.conv_wgrad(x.float(), dy.float(), want, (N, D, H, W, Cin), Cout, k, 0, accumulate, up2=True)
    dw = dw0.to(DEV)
    ops.conv_wgrad(x.to(DEV), dy.to(DEV), dw, (N, D, H, W, Cin), Cout, k, L.ALGO_UMMA, accumulate, up2=True)
    torch.cuda.synchronize()
    assert ops.umma_error_flag() == 0
    got = dw.cpu()
    assert torch.isfinite(got).all()
    err = (got - want).abs().max().item()
    assert err <= 2e-3 * max(1.0, want.abs().max().item()), (case, accumulate, err)


def test_up2_rejected_off_the_tensor_path():
    x = torch.zeros(2 * 2 * 2 * 16, device=DEV)
    w = torch.zeros(16 * 27 * 16, device=DEV)
    y = torch.zeros(8 * 2 * 2 * 2 * 16, device=DEV)
    with pytest.raises(L.R2N2Error):
        L.call('r2n2_conv_fwd', x.data_ptr(), w.data_ptr(), None, None, None, y.data_ptr(), 1, 2, 2, 2, 16, 16,
               3, 3, 3, L.CONV_UP2, L.F32, L.F32, L.ALGO_SIMT, None)

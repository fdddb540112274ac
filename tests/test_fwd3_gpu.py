"""GPU parity of the flat-halo tcgen05 convolution (conv_umma_fwd3.cu: input block loaded once, filter
taps as shifted descriptor views, rolling z-plane ring, streamed weights) against the emulated C-ABI
semantics, bf16 operands.  R2N2_FWD3_FORCE routes small layers to it as well."""
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
    # N, D, H, W, Cin, Cout, k
    (3, 1, 17, 17, 96, 96, (1, 3, 3)),       # conv1b-like: two channel chunks (64 + 32), N = 96
    (2, 1, 40, 70, 96, 128, (1, 3, 3)),      # several x / y blocks, ragged edges
    (2, 1, 33, 33, 128, 128, (1, 3, 3)),
    (2, 1, 9, 9, 64, 64, (1, 3, 3)),
    (1, 16, 16, 16, 32, 32, (3, 3, 3)),      # rolling plane ring, Cin = 32 (SW64)
    (2, 11, 12, 32, 64, 64, (3, 3, 3)),      # two z items (8 + 3 planes)
    (1, 9, 10, 11, 64, 32, (3, 3, 3)),       # conv10a
    (2, 8, 8, 8, 32, 16, (3, 3, 3)),         # conv11: N = 16
    (1, 8, 8, 8, 16, 32, (3, 3, 3)),         # conv11 data gradient: Cin = 16 (SW32)
    (1, 5, 32, 32, 128, 128, (3, 3, 3)),     # two chunks, N = 128, 3-D
    (2, 3, 6, 5, 64, 64, (3, 3, 3)),         # tiny planes (one partial tile)
    (3, 1, 40, 45, 32, 96, (1, 7, 1)),       # row-packed first layer: 7x1 taps, resident weights
    (2, 1, 127, 127, 32, 96, (1, 7, 1)),
]


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('tma_epi,stk,stage', [('0', '0', '1'), ('0', '0', '0'), ('1', '0', '1'), ('0', '1', '1')])
@pytest.mark.parametrize('flags,out_dtype', [(0, torch.float32), (L.EPI_BIAS | L.EPI_LRELU, bf),
                                             (L.EPI_BIAS | L.EPI_RESIDUAL, bf), (L.EPI_RESIDUAL | L.EPI_MASK, bf)])
def test_conv_fwd3(monkeypatch, case, flags, out_dtype, tma_epi, stk, stage):
    monkeypatch.setenv('R2N2_FWD3_FORCE', '1')
    monkeypatch.setenv('R2N2_FWD3_TMA_EPI', tma_epi)      # register epilogue / smem-slab + TMA-store epilogue
    monkeypatch.setenv('R2N2_FWD3_STAGE_EPI', stage)      # warp-staged (8 pixels x 64 B per access) / thread-per-pixel epilogue
    monkeypatch.setenv('R2N2_FWD3_STK', stk)              # one MMA per tap / dx-stacked (N = 3 Cout, shifted sum in the epilogue)
    monkeypatch.setenv('R2N2_FWD3_PAIR', '0')             # single-CTA path (the CTA-pair path: test_conv_fwd3_pair)
    _run_case(case, flags, out_dtype)


def _run_case(case, flags, out_dtype):
    N, D, H, W, Cin, Cout, k = case
    taps = k[0] * k[1] * k[2]
    P = N * D * H * W
    x = _r(P * Cin, 1).to(bf)
    w = _r(Cout * taps * Cin, 2, (2.0 / (taps * Cin)) ** 0.5).to(bf)
    b = _r(Cout, 3)
    res = _r(P * Cout, 4).to(out_dtype)
    mask = _r(P * Cout, 5).to(out_dtype)
    want = torch.empty(P * Cout)
    emu.conv_fwd(x.float(), w.float(), b, mask.float(), res.float(), want, (N, D, H, W, Cin), Cout, k, flags, 0)
    y = torch.full((P * Cout,), float('nan'), dtype=out_dtype, device=DEV)
    ops.conv_fwd(x.to(DEV), w.to(DEV), b.to(DEV), mask.to(DEV), res.to(DEV), y, (N, D, H, W, Cin), Cout, k,
                 flags, L.ALGO_UMMA)
    torch.cuda.synchronize()
    assert ops.umma_error_flag() == 0, 'tcgen05 pipeline timed out'
    got = y.float().cpu()
    assert torch.isfinite(got).all(), 'outputs left unwritten'
    tol = 2e-3 if out_dtype == torch.float32 else 1.2e-2
    err = (got - want).abs().max().item()
    if err > tol * max(1.0, want.abs().max().item()):
        bad = ((got - want).abs() > tol * max(1.0, want.abs().max().item())).nonzero().flatten()
        pix = torch.unique(bad // Cout)[:12].tolist()
        raise AssertionError('fwd3 %s flags %d: max err %.3e, %d bad, first pixels %s' % (case, flags, err, bad.numel(), pix))


PAIR_CASES = [
    (3, 1, 17, 17, 96, 96, (1, 3, 3)),       # conv1b-like, resident half filters; 3 items -> the peer's last item is a ghost
    (2, 1, 127, 127, 96, 96, (1, 3, 3)),     # conv1b itself (two views): T = 2 specialised issue sequence
    (5, 1, 30, 61, 96, 96, (1, 3, 3)),       # ragged blocks, odd number of items
    (2, 1, 40, 70, 96, 128, (1, 3, 3)),      # conv2a-like: streamed half filters (N = 128 across the pair)
    (2, 1, 33, 33, 128, 128, (1, 3, 3)),     # two 64-channel chunks, streamed
    (1, 1, 9, 9, 64, 64, (1, 3, 3)),         # a single item: the peer works on a ghost only
    (1, 16, 16, 16, 32, 32, (3, 3, 3)),      # 3-D rolling plane ring in lockstep
    (2, 11, 12, 32, 64, 64, (3, 3, 3)),      # two z items with different plane counts (8 + 3)
]


@pytest.mark.parametrize('case', PAIR_CASES)
@pytest.mark.parametrize('tma_epi', ['0', '1'])
@pytest.mark.parametrize('flags,out_dtype', [(0, torch.float32), (L.EPI_BIAS | L.EPI_LRELU, bf),
                                             (L.EPI_BIAS | L.EPI_RESIDUAL, bf), (L.EPI_RESIDUAL | L.EPI_MASK, bf)])
def test_conv_fwd3_pair(monkeypatch, case, flags, out_dtype, tma_epi):
    """CTA-pair mode (cluster of two, tcgen05 cta_group::2, M = 256 MMAs, half of the filter rows per CTA,
    multicast commits): same results as the emulated ABI for resident and streamed filters, 2-D and 3-D,
    odd item counts (ghost item) and both epilogues."""
    monkeypatch.setenv('R2N2_FWD3_FORCE', '1')
    monkeypatch.setenv('R2N2_FWD3_TMA_EPI', tma_epi)
    monkeypatch.setenv('R2N2_FWD3_PAIR', '1')
    _run_case(case, flags, out_dtype)


def test_fwd3_matches_per_tap_kernel_bitwise_shape(monkeypatch):
    """Same layer through both tcgen05 kernels (flat-halo vs per-tap): same bf16 products, fp32 accumulation
    in a different order."""
    N, D, H, W, Cin, Cout, k = 2, 6, 9, 13, 64, 64, (3, 3, 3)
    P = N * D * H * W
    x = _r(P * Cin, 1).to(bf).to(DEV)
    w = _r(Cout * 27 * Cin, 2, (2.0 / (27 * Cin)) ** 0.5).to(bf).to(DEV)
    ya = torch.empty(P * Cout, dtype=torch.float32, device=DEV)
    yb = torch.empty(P * Cout, dtype=torch.float32, device=DEV)
    monkeypatch.setenv('R2N2_FWD3_FORCE', '1')
    ops.conv_fwd(x, w, None, None, None, ya, (N, D, H, W, Cin), Cout, k, 0, L.ALGO_UMMA)
    monkeypatch.delenv('R2N2_FWD3_FORCE')
    monkeypatch.setenv('R2N2_UMMA_NO_FWD3', '1')
    ops.conv_fwd(x, w, None, None, None, yb, (N, D, H, W, Cin), Cout, k, 0, L.ALGO_UMMA)
    torch.cuda.synchronize()
    assert ops.umma_error_flag() == 0
    assert (ya - yb).abs().max().item() < 1e-4

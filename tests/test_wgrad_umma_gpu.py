"""GPU parity of the tcgen05 weight-gradient kernel (MN-major operands straight from the
channels-last tensors via TMA) against the emulated C-ABI semantics, bf16 inputs."""
import numpy as np
import pytest
import torch

import r2n2_b200  # noqa: F401
from r2n2_b200 import ops, _lib as L
import emu

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _r(n, seed, scale=1.0):
    return torch.tensor(np.random.RandomState(seed).randn(n) * scale, dtype=torch.float32)


CASES = [
    # N, D, H, W, Cin, Cout, k
    (4, 1, 16, 16, 64, 128, (1, 3, 3)),
    (3, 1, 17, 17, 96, 128, (1, 3, 3)),      # ragged, Cin=96 (64+32 chunk)
    (2, 1, 33, 33, 128, 256, (1, 3, 3)),     # two co tiles
    (2, 1, 9, 9, 256, 256, (1, 3, 3)),       # ci_tile 256, G=2
    (2, 1, 13, 13, 16, 96, (1, 7, 7)),       # conv1a-like: Cin=16 (SW32), 49 taps, Cout=96
    (2, 1, 5, 5, 128, 256, (1, 1, 1)),
    (40, 1, 1, 1, 2304, 1024, (1, 1, 1)),    # fc7 (K = rows of the batch)
    (36, 1, 1, 1, 1024, 512, (1, 1, 1)),     # x-projection slice
    (3, 4, 4, 4, 128, 256, (3, 3, 3)),       # GRU Wh
    (2, 8, 8, 8, 128, 64, (3, 3, 3)),        # Cout=64 (single dy chunk)
    (1, 16, 16, 16, 32, 32, (3, 3, 3)),      # Cin=Cout=32 (SW64 both)
    (1, 8, 8, 8, 128, 64, (1, 1, 1)),
    (2, 6, 5, 7, 64, 64, (3, 3, 3)),
    (1, 8, 8, 8, 64, 32, (3, 3, 3)),         # conv10a-like (stacked taps, TS=2)
    (1, 8, 8, 8, 32, 16, (3, 3, 3)),         # conv11-like (TS=4, Cout=16)
    (1, 8, 8, 8, 16, 32, (3, 3, 3)),         # conv11 data-gradient operand shapes (TS=8)
    (2, 1, 12, 12, 96, 64, (1, 3, 3)),       # Cin=96: no stacking, 32 idle lanes
    (1, 8, 8, 8, 256, 64, (3, 3, 3)),        # two ci tiles
    # flat-halo kernel (conv_umma_wgrad3.cu): Cin, Cout in {16,32,64}, 3x3x3
    (2, 32, 32, 32, 32, 32, (3, 3, 3)),      # conv10b at full spatial size (4 bands of 8 rows)
    (1, 32, 32, 32, 64, 64, (3, 3, 3)),      # conv9b: two MMAs per dy (2 shifts each)
    (1, 5, 9, 33, 64, 16, (3, 3, 3)),        # ragged, dy stacked (8 shifts of 16 channels)
    (3, 4, 4, 4, 16, 16, (3, 3, 3)),
    (1, 3, 70, 11, 32, 64, (3, 3, 3)),       # many bands, last band partial
    # dx-halo mode of the per-tap kernel (box width multiple of 16, one 2-pixel-wider x box per filter row)
    (2, 1, 64, 64, 96, 128, (1, 3, 3)),      # conv2a
    (1, 1, 127, 127, 96, 96, (1, 3, 3)),     # conv1b: box overhangs the image by one column
    (1, 16, 16, 16, 128, 128, (3, 3, 3)),    # conv8b: 9 filter rows
    (2, 1, 30, 40, 128, 128, (1, 3, 3)),     # W = 40 -> 48-wide boxes (17 % overhang)
    (3, 1, 16, 32, 160, 256, (1, 3, 3)),     # two co tiles, ci tile 160
    # kh x 1 filter over one channel chunk (row-packed conv1a): taps stacked along N, one MMA per K16
    (2, 1, 40, 64, 32, 96, (1, 7, 1)),
    (1, 1, 127, 127, 32, 96, (1, 7, 1)),
    (2, 1, 33, 50, 16, 64, (1, 5, 1)),
]


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('accumulate', [False, True])
def test_conv_wgrad_umma(case, accumulate):
    N, D, H, W, Cin, Cout, k = case
    taps = k[0] * k[1] * k[2]
    P = N * D * H * W
    bf = torch.bfloat16
    x, dy = _r(P * Cin, 1).to(bf), _r(P * Cout, 2).to(bf)
    dw0 = _r(Cout * taps * Cin, 3)
    want = dw0.clone()
    emu.conv_wgrad(x.float(), dy.float(), want, (N, D, H, W, Cin), Cout, k, 0, accumulate)
    dw = dw0.to(DEV)
    ops.conv_wgrad(x.to(DEV), dy.to(DEV), dw, (N, D, H, W, Cin), Cout, k, L.ALGO_UMMA, accumulate)
    torch.cuda.synchronize()
    assert ops.umma_error_flag() == 0, 'tcgen05 pipeline timed out'
    got = dw.cpu()
    assert torch.isfinite(got).all()
    err = (got - want).abs().max().item()
    ref = want.abs().max().item()
    tol = 2e-3          # exact bf16 products, fp32 accumulation in a different order
    if err > tol * max(1.0, ref):
        bad = ((got - want).abs() > tol * max(1.0, ref)).nonzero().flatten()
        co = torch.unique(bad // (taps * Cin))[:10].tolist()
        tp = torch.unique((bad // Cin) % taps)[:10].tolist()
        ci = torch.unique(bad % Cin)[:10].tolist()
        raise AssertionError('wgrad umma %s: max err %.3e (ref %.3e) %d bad; co %s taps %s ci %s'
                             % (case, err, ref, bad.numel(), co, tp, ci))

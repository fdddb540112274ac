"""GPU parity of the row-packed first-layer path: r2n2_nchw_to_rowpack + 7x1 conv over 32 packed
channels == the reference 7x7 convolution over RGB (lib/layers.py:304-333)."""
import numpy as np
import pytest
import torch

import r2n2_b200  # noqa: F401
from r2n2_b200 import ops, packing as pk, _lib as L
from oracle import ref_torch as rt
import emu

pytestmark = pytest.mark.gpu


def test_rowpack_layout_and_conv():
    r = np.random.RandomState(0)
    N, C, H, W, Cout = 3, 3, 19, 23, 96
    x = torch.tensor(r.rand(N, C, H, W), dtype=torch.float32)
    w = torch.tensor(r.randn(Cout, C, 7, 7) * 0.1, dtype=torch.float32)
    b = torch.tensor(r.randn(Cout), dtype=torch.float32)
    want_pack = torch.empty(N * H * W * 32)
    emu.nchw_to_rowpack(x, want_pack, 7, 32)
    for dt in (torch.float32, torch.bfloat16):
        y = torch.empty(N * H * W * 32, dtype=dt, device='cuda')
        ops.nchw_to_rowpack(x.cuda(), y, 7, 32)
        assert (y.float().cpu() - want_pack).abs().max() < (1e-7 if dt == torch.float32 else 4e-3)
    # conv through the tensor-core kernel on the packed input vs the oracle's true convolution
    ref = rt.conv2d_same(x, w, b).permute(0, 2, 3, 1).reshape(-1)           # channels-last
    xp = torch.empty(N * H * W * 32, dtype=torch.bfloat16, device='cuda')
    ops.nchw_to_rowpack(x.cuda(), xp, 7, 32)
    wp = pk.pack_conv2d_rowpack(w, 32)
    assert torch.equal(pk.unpack_conv2d_rowpack(wp, 3, 7), w)
    out = torch.empty(N * H * W * Cout, dtype=torch.float32, device='cuda')
    ops.conv_fwd(xp, wp.reshape(-1).to(torch.bfloat16).cuda(), b.cuda(), None, None, out, (N, 1, H, W, 32),
                 Cout, (1, 7, 1), L.EPI_BIAS, L.ALGO_UMMA)
    assert ops.umma_error_flag() == 0
    err = (out.cpu() - ref).abs().max().item()
    assert err < 2e-2 * ref.abs().max().item(), err

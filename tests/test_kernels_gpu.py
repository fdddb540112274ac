"""GPU parity tests, kernel by kernel, through the C ABI (r2n2_b200.ops -> libr2n2_b200.so):
each CUDA kernel against tests/emu.py (the torch-CPU statement of the header's semantics, itself
checked against the oracle on CPU) on seeded inputs, including ragged / edge shapes."""
import numpy as np
import pytest
import torch

import r2n2_b200  # noqa: F401
from r2n2_b200 import ops, _lib as L
import emu

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _r(*shape, seed=0, scale=1.0):
    return torch.tensor(np.random.RandomState(seed).randn(*shape) * scale, dtype=torch.float32)


def _cmp(got, want, tol, what=''):
    got, want = got.float().cpu(), want.float().cpu()
    err = (got - want).abs().max().item()
    ref = want.abs().max().item() + 1e-20
    assert err <= tol * max(1.0, ref), '%s: max err %.3e (ref max %.3e)' % (what, err, ref)


CONV_CASES = [
    # N, D, H, W, Cin, Cout, k
    (2, 1, 9, 7, 4, 96, (1, 7, 7)),       # conv1a-like: Cin padded RGB, 49 taps, ragged M
    (3, 1, 17, 17, 96, 128, (1, 3, 3)),
    (2, 1, 5, 5, 128, 256, (1, 1, 1)),    # 1x1 shortcut
    (5, 1, 1, 1, 2304, 1024, (1, 1, 1)),  # fc7
    (3, 4, 4, 4, 128, 256, (3, 3, 3)),    # GRU hidden conv (u,r fused)
    (1, 8, 8, 8, 32, 2, (3, 3, 3)),       # conv11: Cout=2
    (2, 6, 5, 7, 64, 64, (3, 3, 3)),      # ragged 3-D
    (7, 1, 1, 1, 1024, 8192, (1, 1, 1)),  # x-projection slice
]


@pytest.mark.parametrize('case', CONV_CASES)
@pytest.mark.parametrize('flags', [0, L.EPI_BIAS | L.EPI_LRELU, L.EPI_BIAS | L.EPI_RESIDUAL,
                                   L.EPI_RESIDUAL | L.EPI_MASK])
def test_conv_fwd_simt(case, flags):
    N, D, H, W, Cin, Cout, k = case
    taps = k[0] * k[1] * k[2]
    P = N * D * H * W
    x = _r(P * Cin, seed=1)
    w = _r(Cout * taps * Cin, seed=2, scale=(2.0 / (taps * Cin)) ** 0.5)
    b = _r(Cout, seed=3)
    res = _r(P * Cout, seed=4)
    mask = _r(P * Cout, seed=5)
    want = torch.empty(P * Cout)
    emu.conv_fwd(x, w, b, mask, res, want, (N, D, H, W, Cin), Cout, k, flags, 0)
    y = torch.empty(P * Cout, device=DEV)
    ops.conv_fwd(x.to(DEV), w.to(DEV), b.to(DEV), mask.to(DEV), res.to(DEV), y, (N, D, H, W, Cin),
                 Cout, k, flags, L.ALGO_SIMT)
    _cmp(y, want, 2e-5, 'conv_fwd simt %s flags %d' % (case, flags))


def test_conv_fwd_simt_unaligned_cin():
    """data gradient of conv11: the 'input' is the 2-channel dlogits (Cin=2, K=54)."""
    N, D, H, W, Cin, Cout, k = 1, 8, 8, 8, 2, 32, (3, 3, 3)
    P = N * D * H * W
    x, w = _r(P * Cin, seed=1), _r(Cout * 27 * Cin, seed=2, scale=.2)
    res, mask = _r(P * Cout, seed=4), _r(P * Cout, seed=5)
    for flags in (0, L.EPI_RESIDUAL | L.EPI_MASK):
        want = torch.empty(P * Cout)
        emu.conv_fwd(x, w, None, mask, res, want, (N, D, H, W, Cin), Cout, k, flags, 0)
        y = torch.empty(P * Cout, device=DEV)
        ops.conv_fwd(x.to(DEV), w.to(DEV), None, mask.to(DEV), res.to(DEV), y, (N, D, H, W, Cin), Cout, k,
                     flags, L.ALGO_SIMT)
        _cmp(y, want, 2e-5, 'conv_fwd simt Cin=2')


def test_conv_fwd_in_place_residual():
    """the backward accumulates with residual == y (same buffer)."""
    N, D, H, W, Cin, Cout, k = 2, 1, 6, 6, 32, 32, (1, 3, 3)
    x, w = _r(N * H * W * Cin, seed=1), _r(Cout * 9 * Cin, seed=2, scale=.1)
    acc = _r(N * H * W * Cout, seed=3)
    want = torch.empty_like(acc)
    emu.conv_fwd(x, w, None, None, acc, want, (N, D, H, W, Cin), Cout, k, L.EPI_RESIDUAL, 0)
    buf = acc.to(DEV)
    ops.conv_fwd(x.to(DEV), w.to(DEV), None, None, buf, buf, (N, D, H, W, Cin), Cout, k,
                 L.EPI_RESIDUAL, L.ALGO_SIMT)
    _cmp(buf, want, 2e-5)


@pytest.mark.parametrize('case', CONV_CASES)
@pytest.mark.parametrize('accumulate', [False, True])
def test_conv_wgrad_simt(case, accumulate):
    N, D, H, W, Cin, Cout, k = case
    taps = k[0] * k[1] * k[2]
    P = N * D * H * W
    x, dy = _r(P * Cin, seed=1), _r(P * Cout, seed=2)
    dw0 = _r(Cout * taps * Cin, seed=3)
    want = dw0.clone()
    emu.conv_wgrad(x, dy, want, (N, D, H, W, Cin), Cout, k, 0, accumulate)
    dw = dw0.to(DEV)
    ops.conv_wgrad(x.to(DEV), dy.to(DEV), dw, (N, D, H, W, Cin), Cout, k, L.ALGO_SIMT, accumulate)
    _cmp(dw, want, 5e-5, 'wgrad %s' % (case,))


def test_dgrad_equals_autograd():
    """conv_fwd with r2n2_weight_transpose'd weights is the data gradient."""
    N, D, H, W, Cin, Cout, k = 2, 3, 4, 5, 8, 12, (3, 3, 3)
    x = _r(N, D, H, W, Cin, seed=1).requires_grad_(True)
    w = _r(Cout, 3, 3, 3, Cin, seed=2)
    y = torch.nn.functional.conv3d(x.permute(0, 4, 1, 2, 3), w.permute(0, 4, 1, 2, 3), padding=1)
    dy = _r(*y.shape, seed=3)
    y.backward(dy)
    wt = torch.empty(Cin * 27 * Cout, device=DEV)
    ops.weight_transpose(w.reshape(-1).to(DEV), wt, Cout, 27, Cin)
    dx = torch.empty(N * D * H * W * Cin, device=DEV)
    dyl = dy.permute(0, 2, 3, 4, 1).contiguous().reshape(-1).to(DEV)
    ops.conv_fwd(dyl, wt, None, None, None, dx, (N, D, H, W, Cout), Cin, k, 0, L.ALGO_SIMT)
    _cmp(dx, x.grad.reshape(-1), 2e-5)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_pool_bwd_fused(dtype):
    """r2n2_maxpool2d_bwd_fused: the rectified branch's pre-activation gradient and both bias sums from the
    pool-backward kernel itself == the separate passes (LeakyReLU', column sums) over its outputs."""
    tol = 1e-6 if dtype == torch.float32 else 1e-2
    for (N, H, W, C) in [(2, 64, 64, 96), (3, 33, 33, 128), (2, 9, 9, 256), (5, 127, 127, 96), (1, 3, 3, 8), (2, 17, 17, 40)]:
        a, b = _r(N * H * W * C, seed=1).to(dtype), _r(N * H * W * C, seed=2).to(dtype)
        Ho, Wo = H // 2 + 1, W // 2 + 1
        dy = _r(N * Ho * Wo * C, seed=3).to(dtype)
        for bb, mask in ((b, False), (None, True), (None, False)):
            wdx, wdxb = torch.empty_like(a), (torch.empty_like(a) if bb is not None else None)
            wsa, wsb = torch.full((C,), 0.5), (torch.full((C,), -0.25) if bb is not None else None)
            emu.maxpool_bwd(a, bb, dy, wdx, N, H, W, C, mask, wdxb, wsa, wsb)
            dx = torch.full_like(a, float('nan'), device=DEV)
            dxb = torch.full_like(a, float('nan'), device=DEV) if bb is not None else None
            sa = torch.full((C,), 0.5, device=DEV)                 # the sums ACCUMULATE into the buffers
            sb = torch.full((C,), -0.25, device=DEV) if bb is not None else None
            ops.maxpool_bwd(a.to(DEV), None if bb is None else bb.to(DEV), dy.to(DEV), dx, N, H, W, C, mask,
                            dxb, sa, sb)
            _cmp(dx, wdx, tol, 'fused pool bwd dx')
            scale = max(1.0, float(wsa.abs().max()))
            assert (sa.cpu() - wsa).abs().max().item() <= 2e-3 * scale, 'colsum_a'
            if bb is not None:
                _cmp(dxb, wdxb, tol, 'fused pool bwd dxb')
                assert (sb.cpu() - wsb).abs().max().item() <= 2e-3 * max(1.0, float(wsb.abs().max())), 'colsum_b'
            # only the column sum of a (no second output)
            dx2 = torch.empty_like(dx)
            sa2 = torch.zeros(C, device=DEV)
            ops.maxpool_bwd(a.to(DEV), None if bb is None else bb.to(DEV), dy.to(DEV), dx2, N, H, W, C, mask,
                            None, sa2, None)
            assert torch.equal(dx2, dx)
            assert (sa2.cpu() + 0.5 - wsa).abs().max().item() <= 2e-3 * scale
    with pytest.raises(RuntimeError):
        ops.maxpool_bwd(a.to(DEV), None, dy.to(DEV), dx, N, H, W, C, False, torch.empty_like(dx), None, None)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_pool_codes(dtype):
    """r2n2_maxpool2d_fwd_code / _bwd_code: the backward pass from dy + one code byte per output element (arg-max
    positions incl. ties, sign of the rectified operand) is BIT-IDENTICAL to r2n2_maxpool2d_bwd_fused on a, b, dy;
    the code bytes equal the emulated ones.  Values are quantised so that ties do occur."""
    for (N, H, W, C) in [(2, 64, 64, 96), (3, 33, 33, 128), (2, 9, 9, 256), (5, 127, 127, 96), (1, 3, 3, 8), (2, 17, 17, 40),
                         (1, 4, 6, 512)]:
        q = lambda t: (t * 4).round() / 4           # coarse grid: many exact ties and exact zeros
        a, b = q(_r(N * H * W * C, seed=1)).to(dtype), q(_r(N * H * W * C, seed=2)).to(dtype)
        Ho, Wo = H // 2 + 1, W // 2 + 1
        dy = _r(N * Ho * Wo * C, seed=3).to(dtype)
        for bb, mask in ((b, False), (None, True), (None, False)):
            y0 = torch.empty(N * Ho * Wo * C, dtype=dtype, device=DEV)
            y1 = torch.empty_like(y0)
            code = torch.full((N * Ho * Wo * C,), 255, dtype=torch.uint8, device=DEV)
            ad, bd = a.to(DEV), None if bb is None else bb.to(DEV)
            ops.maxpool_fwd(ad, bd, y0, N, H, W, C)
            ops.maxpool_fwd(ad, bd, y1, N, H, W, C, code)
            assert torch.equal(y0, y1)
            wy, wcode = torch.empty(N * Ho * Wo * C, dtype=dtype), torch.empty(N * Ho * Wo * C, dtype=torch.uint8)
            emu.maxpool_fwd(a, bb, wy, N, H, W, C, wcode)
            assert torch.equal(code.cpu(), wcode), 'code bytes'
            for with_b_out, with_sums in ((True, True), (True, False), (False, True), (False, False)):
                if with_b_out and bb is None:
                    continue
                dx0, dx1 = torch.full_like(ad, float('nan')), torch.full_like(ad, float('nan'))
                dxb0 = torch.full_like(ad, float('nan')) if with_b_out else None
                dxb1 = torch.full_like(ad, float('nan')) if with_b_out else None
                sa0 = torch.full((C,), 0.5, device=DEV) if with_sums else None
                sa1 = torch.full((C,), 0.5, device=DEV) if with_sums else None
                sb0 = torch.zeros(C, device=DEV) if (with_sums and with_b_out) else None
                sb1 = torch.zeros(C, device=DEV) if (with_sums and with_b_out) else None
                ops.maxpool_bwd(ad, bd, dy.to(DEV), dx0, N, H, W, C, mask, dxb0, sa0, sb0)
                ops.maxpool_bwd_code(code, dy.to(DEV), dx1, N, H, W, C, mask, dxb1, sa1, sb1)
                assert torch.equal(dx0, dx1), 'dx from codes'
                if with_b_out:
                    assert torch.equal(dxb0, dxb1), 'dxb from codes'
                if with_sums:
                    assert (sa0 - sa1).abs().max().item() <= 1e-3 * max(1.0, float(sa0.abs().max()))
                    if with_b_out:
                        assert (sb0 - sb1).abs().max().item() <= 1e-3 * max(1.0, float(sb0.abs().max()))
    with pytest.raises(RuntimeError):       # one sign per position: lrelu_mask and dxb together are refused
        ops.maxpool_bwd_code(code, dy.to(DEV), dx1, N, H, W, C, True, torch.empty_like(dx1), None, None)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_pool_unpool(dtype):
    tol = 1e-6 if dtype == torch.float32 else 1e-2
    for (N, H, W, C) in [(2, 127, 127, 8), (3, 33, 33, 16), (2, 5, 5, 256), (1, 3, 3, 4), (1, 64, 64, 96)]:
        if dtype == torch.bfloat16 and C % 8:
            continue                      # 16-byte channel vectors: 8 x bf16
        a, b = _r(N * H * W * C, seed=1).to(dtype), _r(N * H * W * C, seed=2).to(dtype)
        Ho, Wo = H // 2 + 1, W // 2 + 1
        dy = _r(N * Ho * Wo * C, seed=3).to(dtype)
        for bb in (None, b):
            want, y = torch.empty(N * Ho * Wo * C, dtype=dtype), torch.empty(N * Ho * Wo * C, dtype=dtype, device=DEV)
            emu.maxpool_fwd(a, bb, want, N, H, W, C)
            ops.maxpool_fwd(a.to(DEV), None if bb is None else bb.to(DEV), y, N, H, W, C)
            _cmp(y, want, tol, 'pool fwd')
            for mask in (False, True):
                wdx, dx = torch.empty_like(a), torch.empty_like(a, device=DEV)
                emu.maxpool_bwd(a, bb, dy, wdx, N, H, W, C, mask)
                ops.maxpool_bwd(a.to(DEV), None if bb is None else bb.to(DEV), dy.to(DEV), dx, N, H, W, C, mask)
                _cmp(dx, wdx, tol, 'pool bwd')
    for (B, D, H, W, C) in [(2, 4, 4, 4, 128), (1, 3, 5, 2, 8)]:
        a, b = _r(B * D * H * W * C, seed=1).to(dtype), _r(B * D * H * W * C, seed=2).to(dtype)
        for bb in (None, b):
            want, y = torch.empty(a.numel() * 8, dtype=dtype), torch.empty(a.numel() * 8, dtype=dtype, device=DEV)
            emu.unpool_fwd(a, bb, want, B, D, H, W, C)
            ops.unpool_fwd(a.to(DEV), None if bb is None else bb.to(DEV), y, B, D, H, W, C)
            _cmp(y, want, tol, 'unpool fwd')
        dy = _r(a.numel() * 8, seed=5).to(dtype)
        wdx, dx = torch.empty_like(a), torch.empty_like(a, device=DEV)
        emu.unpool_bwd(dy, wdx, B, D, H, W, C)
        ops.unpool_bwd(dy.to(DEV), dx, B, D, H, W, C)
        _cmp(dx, wdx, 0.0 if dtype == torch.float32 else tol, 'unpool bwd')


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_lrelu_bwd_colsum(dtype):
    """r2n2_lrelu_bwd_colsum: dx bit-identical to r2n2_lrelu_bwd, column sums of the stored dx accumulate into fp32."""
    for rows, C in [(1237, 32), (36 * 64, 128), (5000, 64), (7, 8), (20011, 40), (333, 512)]:
        a, g, ad = _r(rows * C, seed=1).to(dtype), _r(rows * C, seed=2).to(dtype), _r(rows * C, seed=3).to(dtype)
        for addend in (None, ad):
            dx0 = torch.empty(rows * C, dtype=dtype, device=DEV)
            dx1 = torch.full((rows * C,), float('nan'), dtype=dtype, device=DEV)
            cs = torch.full((C,), 0.25, device=DEV)
            add_d = None if addend is None else addend.to(DEV)
            ops.lrelu_bwd(a.to(DEV), g.to(DEV), add_d, dx0)
            ops.lrelu_bwd(a.to(DEV), g.to(DEV), add_d, dx1, colsum=cs, C=C)
            assert torch.equal(dx0, dx1)
            want = dx0.double().view(rows, C).sum(0).cpu() + 0.25
            assert (cs.cpu().double() - want).abs().max().item() <= 1e-4 * max(1.0, float(want.abs().max()))
            wdx, wcs = torch.empty(rows * C, dtype=dtype), torch.full((C,), 0.25)
            emu.lrelu_bwd(a, g, addend, wdx, wcs, C)
            _cmp(dx1, wdx, 1e-6 if dtype == torch.float32 else 1e-2, 'lrelu_bwd_colsum')


def test_pointwise_and_layout():
    n = 4 * 1237
    a, b = _r(n, seed=1), _r(n, seed=2)
    for name, args in [('add', (a, b)), ('lrelu_fwd', (a,)), ('lrelu_bwd', (a, b, None)),
                       ('lrelu_bwd', (a, b, a))]:
        want, y = torch.empty(n), torch.empty(n, device=DEV)
        getattr(emu, name)(*args, want)
        getattr(ops, name)(*[None if t is None else t.to(DEV) for t in args], y)
        _cmp(y, want, 1e-6, name)
    for op in range(7):       # forward ops 0..3, SIGMOID_BWD / TANH_BWD / NEG (lib/layers.py:649-676 under tensor.grad)
        want, y = torch.empty(n), torch.empty(n, device=DEV)
        emu.eltwise(op, a, b, want)
        ops.eltwise(op, a.to(DEV), b.to(DEV), y)
        _cmp(y, want, 2e-6, 'eltwise %d' % op)
    x = _r(3, 3, 11, 13, seed=4)
    for cpad, wpad, xoff in [(4, None, 0), (16, None, 0), (8, 24, 3)]:
        wp = wpad or 13
        want, y = torch.empty(3 * 11 * wp * cpad), torch.empty(3 * 11 * wp * cpad, device=DEV)
        emu.nchw_to_nhwc(x, want, cpad, wpad, xoff)
        ops.nchw_to_nhwc(x.to(DEV), y, cpad, wpad, xoff)
        _cmp(y, want, 0.0, 'nchw_to_nhwc')
    w = _r(10, 27, 12, seed=5).reshape(-1)
    want, wt = torch.empty(w.numel()), torch.empty(w.numel(), device=DEV)
    emu.weight_transpose(w, want, 10, 27, 12)
    ops.weight_transpose(w.to(DEV), wt, 10, 27, 12)
    _cmp(wt, want, 0.0, 'weight_transpose')
    # every weight tensor of a store in one launch
    shapes = [(10, 27, 12), (96, 9, 96), (33, 1, 70), (128, 27, 16), (200, 1, 1024)]
    flat, rows, dst, tiles = [], [], 0, 0
    for i, (co, tp, ci) in enumerate(shapes):
        t_ = _r(co * tp * ci, seed=20 + i)
        rows.append([sum(x.numel() for x in flat), dst, co, tp, ci, tiles])
        flat.append(t_)
        dst += co * tp * ci
        tiles += ((ci + 31) // 32) * ((co + 63) // 64) * tp
    base = torch.cat(flat).to(DEV)
    for dt in (torch.float32, torch.bfloat16):
        out = torch.full((dst,), float('nan'), dtype=dt, device=DEV)
        table = torch.tensor(rows, dtype=torch.int64, device=DEV)
        L.call('r2n2_weight_transpose_batch', base.data_ptr(), out.data_ptr(), table.data_ptr(), len(rows), tiles,
               L.F32 if dt == torch.float32 else L.BF16, None)
        for (co, tp, ci), r_, t_ in zip(shapes, rows, flat):
            want = torch.empty(t_.numel())
            emu.weight_transpose(t_, want, co, tp, ci)
            _cmp(out[r_[1]:r_[1] + t_.numel()], want.to(dt), 0.0, 'weight_transpose_batch')
    # the same from a bf16 mirror (64 x 64 tiles; even offsets and channel counts): bit-identical to the fp32-source path
    shapes = [(10, 27, 12), (96, 9, 96), (34, 1, 70), (128, 27, 16), (200, 1, 1024), (2, 27, 32), (64, 1, 64)]
    flat, rows, dst, tiles = [], [], 0, 0
    for i, (co, tp, ci) in enumerate(shapes):
        t_ = _r(co * tp * ci, seed=40 + i)
        rows.append([sum(x.numel() for x in flat), dst, co, tp, ci, tiles])
        flat.append(t_)
        dst += (co * tp * ci + 63) // 64 * 64
        tiles += ((ci + 63) // 64) * ((co + 63) // 64) * tp
    mirror = torch.cat(flat).to(torch.bfloat16).to(DEV)
    out = torch.full((dst,), float('nan'), dtype=torch.bfloat16, device=DEV)
    table = torch.tensor(rows, dtype=torch.int64, device=DEV)
    L.call('r2n2_weight_transpose_batch_bf16', mirror.data_ptr(), out.data_ptr(), table.data_ptr(), len(rows), tiles, None)
    for (co, tp, ci), r_, t_ in zip(shapes, rows, flat):
        want = torch.empty(t_.numel())
        emu.weight_transpose(t_.to(torch.bfloat16).float(), want, co, tp, ci)
        got = out[r_[1]:r_[1] + t_.numel()].float().cpu()
        assert torch.equal(got, want), 'weight_transpose_batch_bf16 %s' % ((co, tp, ci),)
    # flat kernel (C/vector <= 256 vectors per row), ragged row counts, and the generic kernel
    for rows, C in [(1000, 96), (77, 2), (3, 1024), (5000, 256), (40003, 32), (12345, 64), (999, 16),
                    (2, 4096), (70001, 128)]:
        for dt in (torch.float32, torch.bfloat16):
            dy = _r(rows * C, seed=6).to(dt)
            for acc in (False, True):
                want = _r(C, seed=7)
                db = want.to(DEV)
                emu.bias_grad(dy.float(), want, C, acc)
                ops.bias_grad(dy.to(DEV), db, C, acc)
                _cmp(db, want, 2e-5, 'bias_grad %d x %d %s' % (rows, C, dt))


def test_gate_kernels():
    B = 3
    hn = B * 64 * 128
    t = lambda n, s: _r(n, seed=s, scale=.7)
    conv_ur, xp_ur, b_ur, h = t(2 * hn, 1), t(2 * hn, 2), t(256, 3), t(hn, 4)
    for cv, hh in ((conv_ur, h), (None, None)):
        outs_w = [torch.empty(hn) for _ in range(3)]
        outs_g = [torch.empty(hn, device=DEV) for _ in range(3)]
        emu.gru_ur_fwd(cv, xp_ur, b_ur, hh, *outs_w, B)
        ops.gru_ur_fwd(None if cv is None else cv.to(DEV), xp_ur.to(DEV), b_ur.to(DEV),
                       None if hh is None else hh.to(DEV), *outs_g, B)
        for g, w in zip(outs_g, outs_w):
            _cmp(g, w, 2e-6, 'gru_ur_fwd')
    u = torch.sigmoid(t(hn, 5))
    conv_c, xp_c, b_c = t(hn, 6), t(hn, 7), t(128, 8)
    ow = [torch.empty(hn), torch.empty(hn)]
    og = [torch.empty(hn, device=DEV), torch.empty(hn, device=DEV)]
    emu.gru_out_fwd(conv_c, xp_c, b_c, u, h, *ow, B)
    ops.gru_out_fwd(conv_c.to(DEV), xp_c.to(DEV), b_c.to(DEV), u.to(DEV), h.to(DEV), *og, B)
    for g, w in zip(og, ow):
        _cmp(g, w, 2e-6, 'gru_out_fwd')
    dh, c, r = t(hn, 9), torch.tanh(t(hn, 10)), torch.sigmoid(t(hn, 11))
    for hh in (h, None):
        w_ur, w_c, w_p = torch.zeros(2 * hn), torch.empty(hn), torch.empty(hn)
        g_ur, g_c, g_p = torch.zeros(2 * hn, device=DEV), torch.empty(hn, device=DEV), torch.empty(hn, device=DEV)
        emu.gru_out_bwd(dh, u, c, hh, w_ur, w_c, w_p, B)
        ops.gru_out_bwd(dh.to(DEV), u.to(DEV), c.to(DEV), None if hh is None else hh.to(DEV), g_ur, g_c, g_p, B)
        d_rh = t(hn, 12)
        emu.gru_r_bwd(d_rh, h, r, w_ur, w_p, B)
        ops.gru_r_bwd(d_rh.to(DEV), h.to(DEV), r.to(DEV), g_ur, g_p, B)
        for g, w in ((g_ur, w_ur), (g_c, w_c), (g_p, w_p)):
            _cmp(g, w, 2e-6, 'gru bwd')
    # LSTM
    conv, xp, bb, s = t(3 * hn, 13), t(3 * hn, 14), t(384, 15), t(hn, 16)
    for cv, ss in ((conv, s), (None, None)):
        ow = [torch.empty(3 * hn), torch.empty(hn), torch.empty(hn)]
        og = [torch.empty(3 * hn, device=DEV), torch.empty(hn, device=DEV), torch.empty(hn, device=DEV)]
        emu.lstm_fwd(cv, xp, bb, ss, *ow, B)
        ops.lstm_fwd(None if cv is None else cv.to(DEV), xp.to(DEV), bb.to(DEV), None if ss is None else ss.to(DEV), *og, B)
        for g, w in zip(og, ow):
            _cmp(g, w, 2e-6, 'lstm_fwd')
        ds = t(hn, 17)
        for dsi in (ds, None):
            bw = [torch.empty(3 * hn), torch.empty(hn)]
            bg = [torch.empty(3 * hn, device=DEV), torch.empty(hn, device=DEV)]
            emu.lstm_bwd(dh, dsi, ow[0], ss, ow[2], *bw, B)
            ops.lstm_bwd(dh.to(DEV), None if dsi is None else dsi.to(DEV), og[0], None if ss is None else ss.to(DEV), og[2], *bg, B)
            for g, w in zip(bg, bw):
                _cmp(g, w, 3e-6, 'lstm_bwd')


def test_softmax_ce_adam_voxel():
    B, nv = 3, 32
    logits = _r(B * nv ** 3 * 2, seed=1, scale=2.0)
    occ = np.random.RandomState(2).rand(B, nv, nv, nv) < 0.1
    y = torch.zeros(B, nv, 2, nv, nv)
    y[:, :, 1] = torch.tensor(occ, dtype=torch.float32)
    y[:, :, 0] = 1 - y[:, :, 1]
    for yy in (y, None):
        wp, wl, wd = torch.empty(B, nv, 2, nv, nv), torch.zeros(1), torch.empty(logits.numel())
        gp, gl, gd = torch.empty(B, nv, 2, nv, nv, device=DEV), torch.zeros(1, device=DEV), torch.empty(logits.numel(), device=DEV)
        emu.softmax_ce3d(logits, yy, wp, wl, wd, B, nv, 0.125)
        ops.softmax_ce3d(logits.to(DEV), None if yy is None else yy.to(DEV), gp, gl, gd, B, nv, 0.125)
        _cmp(gp, wp, 2e-6, 'softmax pred')
        _cmp(gd, wd, 2e-6, 'softmax dlogits')
        assert abs(gl.item() - wl.item()) < 2e-5 * abs(wl.item())
    # padded logits (conv11 on the tensor path: 16 channels per voxel), bf16 gradient rows
    lp = torch.zeros(B * nv ** 3, 16)
    lp[:, :2] = logits.view(-1, 2)
    lp[:, 2:] = 7.0                                   # must be ignored
    wd16 = torch.empty(lp.numel(), dtype=torch.bfloat16)
    gd16 = torch.full((lp.numel(),), float('nan'), dtype=torch.bfloat16, device=DEV)
    wl16, gl16 = torch.zeros(1), torch.zeros(1, device=DEV)
    emu.softmax_ce3d(lp.reshape(-1), y, wp, wl16, wd16, B, nv, 0.125, 16)
    ops.softmax_ce3d(lp.reshape(-1).to(DEV), y.to(DEV), gp, gl16, gd16, B, nv, 0.125, 16)
    _cmp(gp, wp, 2e-6, 'softmax pred (padded)')
    _cmp(gd16, wd16, 4e-3, 'softmax dlogits (padded, bf16)')
    assert abs(gl16.item() - wl16.item()) < 2e-5 * abs(wl16.item())
    # prediction-only (inference) call
    gp2 = torch.empty(B, nv, 2, nv, nv, device=DEV)
    ops.softmax_ce3d(logits.to(DEV), None, gp2, None, None, B, nv, 1.0)
    _cmp(gp2, wp, 2e-6)
    # Adam / SGD: ragged n, decay boundary inside a float4
    n, nd = 10007, 5002
    p, g = _r(n, seed=3), _r(n, seed=4)
    wp_, wm, wv = p.clone(), torch.zeros(n), torch.zeros(n)
    gp_, gm, gv = p.to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    mir = torch.empty(n, dtype=torch.bfloat16, device=DEV)
    for step in range(1, 4):
        lr_t = 1e-2 * (1 - 0.999 ** step) ** 0.5 / (1 - 0.9 ** step)
        emu.adam(wp_, g, wm, wv, None, nd, lr_t, 0.9, 0.999, 1e-8, 5e-5, 0.5)
        ops.adam(gp_, g.to(DEV), gm, gv, mir, nd, lr_t, 0.9, 0.999, 1e-8, 5e-5, 0.5)
    _cmp(gp_, wp_, 2e-6, 'adam p'); _cmp(gm, wm, 2e-6, 'adam m'); _cmp(gv, wv, 2e-6, 'adam v')
    _cmp(mir, wp_, 1e-2, 'adam mirror')
    wv2, gv2 = torch.zeros(n), torch.zeros(n, device=DEV)
    emu.sgd(wp_, g, wv2, None, nd, 0.1, 0.9, 5e-5, 1.0)
    ops.sgd(gp_, g.to(DEV), gv2, None, nd, 0.1, 0.9, 5e-5, 1.0)
    _cmp(gp_, wp_, 3e-6, 'sgd')
    # voxel IoU counts against the golden fixture generated by the real lib/voxel.py
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'voxel_eval.npz'))
    p1 = torch.tensor(z['preds1'].astype(np.float32))
    oc = torch.tensor(np.unpackbits(z['occ'])[:p1.numel()].reshape(p1.shape).astype(np.float32))
    pred = torch.stack([1 - p1, p1], dim=2).contiguous()
    gt = torch.stack([1 - oc, oc], dim=2).contiguous()
    for ti, th in enumerate(z['thresholds']):
        counts = ops.voxel_eval(pred.to(DEV), gt.to(DEV), float(th)).cpu().numpy()
        assert np.array_equal(counts, z['counts'][:, ti]), th


def test_errors_are_loud():
    x = torch.zeros(16, device=DEV)
    with pytest.raises(L.R2N2Error):       # wgrad needs Cin % 4 == 0
        ops.conv_wgrad(torch.zeros(2 * 3, device=DEV), torch.zeros(2 * 4, device=DEV),
                       torch.zeros(4 * 3, device=DEV), (2, 1, 1, 1, 3), 4, (1, 1, 1), L.ALGO_SIMT, False)
    with pytest.raises(L.R2N2Error):       # even kernel extent
        ops.conv_fwd(x, torch.zeros(64, device=DEV), None, None, None, x, (1, 1, 2, 2, 4), 4, (1, 2, 2), 0,
                     L.ALGO_SIMT)
    with pytest.raises(RuntimeError):      # CPU tensor
        ops.add(torch.zeros(4), torch.zeros(4), torch.zeros(4))


# ------------------------------------------------------------------ tcgen05 / TMA kernel
UMMA_CASES = [
    # N, D, H, W, Cin, Cout, k
    (2, 1, 16, 16, 64, 128, (1, 3, 3)),     # exact tiles, one K block per tap
    (3, 1, 17, 17, 96, 128, (1, 3, 3)),     # ragged box, partial K block (96 = 64 + 32)
    (2, 1, 33, 33, 128, 256, (1, 3, 3)),    # two N tiles, two K blocks per tap
    (2, 1, 9, 9, 16, 96, (1, 7, 7)),        # conv1a-like: Cin=16 (SW32), 49 taps, BN=96
    (2, 1, 5, 5, 256, 256, (1, 1, 1)),      # 1x1
    (5, 1, 1, 1, 2304, 1024, (1, 1, 1)),    # fc7: M tile mostly out of bounds
    (3, 4, 4, 4, 128, 256, (3, 3, 3)),      # GRU hidden conv, box spans 2 samples
    (2, 8, 8, 8, 128, 128, (3, 3, 3)),
    (1, 16, 16, 16, 32, 32, (3, 3, 3)),     # Cin=32 (SW64), BN=32
    (1, 8, 8, 8, 128, 64, (1, 1, 1)),       # conv9c
    (180, 1, 1, 1, 1024, 512, (1, 1, 1)),   # x-projection slice
    (2, 6, 5, 7, 64, 64, (3, 3, 3)),        # ragged 3-D
    # the single-wave layers that run on MULTICAST CLUSTERS (each CTA loads 1/C of every A box for the whole cluster)
    (36, 4, 4, 4, 128, 256, (3, 3, 3)),     # GRU hidden conv at batch 36: 18 M tiles, slices = planes of one image
    (36, 4, 4, 4, 256, 128, (3, 3, 3)),     # its data gradient (u/r gates): K = 6912
    (32, 4, 4, 4, 128, 128, (3, 3, 3)),     # candidate conv at batch 32 (config 5)
    (180, 1, 1, 1, 2304, 1024, (1, 1, 1)),  # fc7 at T*B = 180: second M tile mostly out of bounds, slices = images
    (180, 1, 1, 1, 1024, 2304, (1, 1, 1)),  # fc7 data gradient
    (180, 1, 5, 5, 256, 256, (1, 3, 3)),    # conv6a: 125-row boxes (no aligned slices -> no cluster)
    (7, 4, 4, 4, 128, 384, (3, 3, 3)),      # LSTM hidden conv, odd batch (last box half out of bounds)
]


@pytest.mark.parametrize('case', UMMA_CASES)
@pytest.mark.parametrize('flags,out_dtype', [(0, torch.float32), (L.EPI_BIAS | L.EPI_LRELU, torch.bfloat16),
                                             (L.EPI_BIAS | L.EPI_RESIDUAL, torch.bfloat16),
                                             (L.EPI_RESIDUAL | L.EPI_MASK, torch.bfloat16)])
@pytest.mark.parametrize('mc', ['0', '1'])
def test_conv_fwd_umma_bf16(monkeypatch, case, flags, out_dtype, mc):
    if mc == '1' and case[0] * case[1] * case[2] * case[3] > 128 * 148:
        pytest.skip('multicast clusters only exist for single-wave layers')
    monkeypatch.setenv('R2N2_UMMA_MC', mc)        # multicast clusters (opt-in) for the single-wave layers
    N, D, H, W, Cin, Cout, k = case
    taps = k[0] * k[1] * k[2]
    P = N * D * H * W
    bf = torch.bfloat16
    x = _r(P * Cin, seed=1).to(bf)
    w = _r(Cout * taps * Cin, seed=2, scale=(2.0 / (taps * Cin)) ** 0.5).to(bf)
    b = _r(Cout, seed=3)
    res = _r(P * Cout, seed=4).to(out_dtype)
    mask = _r(P * Cout, seed=5).to(out_dtype)
    want = torch.empty(P * Cout, dtype=torch.float32)
    emu.conv_fwd(x.float(), w.float(), b, mask.float(), res.float(), want, (N, D, H, W, Cin), Cout, k, flags, 0)
    y = torch.full((P * Cout,), float('nan'), dtype=out_dtype, device=DEV)
    ops.conv_fwd(x.to(DEV), w.to(DEV), b.to(DEV), mask.to(DEV), res.to(DEV), y, (N, D, H, W, Cin), Cout, k,
                 flags, L.ALGO_UMMA)
    torch.cuda.synchronize()
    assert ops.umma_error_flag() == 0, 'tcgen05 pipeline timed out'
    got = y.float().cpu()
    assert torch.isfinite(got).all(), 'unwritten / non-finite outputs: %d' % (~torch.isfinite(got)).sum().item()
    tol = 2e-3 if out_dtype == torch.float32 else 1.2e-2       # exact bf16 products, fp32 accumulate
    err = (got - want).abs().max().item()
    ref = want.abs().max().item()
    if err > tol * max(1.0, ref):
        bad = ((got - want).abs() > tol * max(1.0, ref)).nonzero().flatten()
        rows = torch.unique(bad // Cout)[:12].tolist()
        cols = torch.unique(bad % Cout)[:12].tolist()
        raise AssertionError('umma %s flags %d: max err %.3e (ref %.3e), %d bad; rows %s cols %s'
                             % (case, flags, err, ref, bad.numel(), rows, cols))


@pytest.mark.parametrize('case', [(2, 1, 16, 16, 64, 128, (1, 3, 3)), (3, 1, 17, 17, 96, 96, (1, 3, 3)),
                                  (2, 6, 5, 7, 64, 64, (3, 3, 3)), (1, 16, 16, 16, 32, 32, (3, 3, 3)),
                                  (2, 1, 5, 5, 256, 256, (1, 1, 1)), (2, 1, 9, 9, 64, 384, (1, 3, 3))])
@pytest.mark.parametrize('flags', [0, L.EPI_RESIDUAL | L.EPI_MASK])
def test_conv_fwd_colsum(case, flags):
    """r2n2_conv_fwd_colsum: column sums of the stored (bf16) output accumulate into an fp32 vector -- fused
    in the TMA epilogue, or a second pass (flat-halo kernel, several N tiles)."""
    N, D, H, W, Cin, Cout, k = case
    taps = k[0] * k[1] * k[2]
    P = N * D * H * W
    bf = torch.bfloat16
    x = _r(P * Cin, seed=1).to(bf)
    w = _r(Cout * taps * Cin, seed=2, scale=(2.0 / (taps * Cin)) ** 0.5).to(bf)
    res = _r(P * Cout, seed=4).to(bf)
    mask = _r(P * Cout, seed=5).to(bf)
    want = torch.empty(P * Cout, dtype=bf)
    cs0 = _r(Cout, seed=6)
    want_cs = cs0.clone()
    emu.conv_fwd(x.float(), w.float(), None, mask.float(), res.float(), want, (N, D, H, W, Cin), Cout, k, flags, 0,
                 colsum=want_cs)
    y = torch.full((P * Cout,), float('nan'), dtype=bf, device=DEV)
    cs = cs0.to(DEV)
    ops.conv_fwd(x.to(DEV), w.to(DEV), None, mask.to(DEV), res.to(DEV), y, (N, D, H, W, Cin), Cout, k, flags,
                 L.ALGO_UMMA, colsum=cs)
    torch.cuda.synchronize()
    assert ops.umma_error_flag() == 0
    _cmp(y, want, 1.2e-2, 'conv_fwd_colsum y')
    # the sums are over the bf16 values actually stored: compare with the sums of OUR output, tightly
    own = cs0 + y.float().cpu().view(-1, Cout).double().sum(0).float()
    assert (cs.cpu() - own).abs().max().item() <= 2e-3 * max(1.0, own.abs().max().item())
    assert (cs.cpu() - want_cs).abs().max().item() <= 5e-2 * max(1.0, want_cs.abs().max().item())


def test_average_precision_golden_and_oracle():
    """r2n2_average_precision (shared-memory bitonic sort + scans, one CTA per sample) against scikit-learn's
    outputs (golden fixture) and, on other sizes / tie patterns, against the oracle restatement."""
    import os
    from oracle import ref_numpy as rn
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'average_precision.npz'))
    p1 = torch.tensor(z['p1'])
    oc = torch.tensor(np.unpackbits(z['occ'])[:p1.numel()].reshape(p1.shape).astype(np.float32))
    pred = torch.stack([1 - p1, p1], dim=2).contiguous()
    gt = torch.stack([1 - oc, oc], dim=2).contiguous()
    ap = ops.average_precision(pred.to(DEV), gt.to(DEV)).cpu().numpy()
    assert np.abs(ap - z['ap']).max() <= 1e-12, (ap, z['ap'])
    r = np.random.RandomState(5)
    for nv, B, levels in [(8, 3, 0), (16, 2, 5), (32, 2, 2), (20, 2, 11)]:
        p = r.rand(B, nv, nv, nv).astype(np.float32)
        if levels:
            p = (np.round(p * levels) / levels).astype(np.float32)
        o = (r.rand(B, nv, nv, nv) < 0.2).astype(np.float32)
        pr = torch.tensor(np.stack([1 - p, p], axis=2))
        g = torch.tensor(np.stack([1 - o, o], axis=2))
        got = ops.average_precision(pr.to(DEV).contiguous(), g.to(DEV).contiguous()).cpu().numpy()
        want = np.array([rn.average_precision(o[j], p[j]) for j in range(B)])
        assert np.abs(got - want).max() <= 1e-12, (nv, levels, got, want)


def test_preprocess_views_golden(monkeypatch):
    """r2n2_preprocess_views against the outputs of the REAL lib/data_augmentation.preprocess_img (golden
    fixture; bit-exact fp32), the host mirror's random draws against the reference's (same seed -> same crop /
    flip / background), and other sizes against the oracle."""
    import os
    from oracle import ref_numpy as rn
    from r2n2_b200.lib import data_augmentation as da
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'preprocess.npz'))
    out = da.preprocess_views(z['rgba'], params=z['params']).cpu().numpy()          # (n,3,127,127)
    assert np.array_equal(out.transpose(0, 2, 3, 1), z['out'])
    for i in range(z['rgba'].shape[0]):                                              # reference-signature wrapper
        np.random.seed(int(z['seeds'][i]))
        got = da.preprocess_img(z['rgba'][i], train=bool(z['train'][i]))
        assert np.array_equal(got, z['out'][i]), i
    np.random.seed(int(z['seeds'][0]))
    assert np.array_equal(da.draw_params(1, True)[0], z['params'][0])
    r = np.random.RandomState(3)
    rgba = r.randint(0, 256, (5, 40, 52, 4)).astype(np.uint8)
    rgba[..., 3] = np.where(r.rand(5, 40, 52) < 0.5, 0, rgba[..., 3])
    prm = np.array([[r.randint(0, 40 - 31 + 1), r.randint(0, 52 - 33 + 1), r.randint(0, 2), 225, 240, 255] for _ in range(5)],
                   np.int32)
    o = torch.empty(5, 3, 31, 33, device=DEV)
    ops.preprocess_views(torch.tensor(rgba).to(DEV), torch.tensor(prm).to(DEV), o)
    for i in range(5):
        want = rn.preprocess_img(rgba[i], prm[i], (31, 33))
        assert np.array_equal(o[i].permute(1, 2, 0).cpu().numpy(), want), i
    with pytest.raises(ValueError):
        da.preprocess_views(np.zeros((1, 100, 100, 4), np.uint8))                    # crop window does not fit

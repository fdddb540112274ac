"""TEST INFRASTRUCTURE: torch-CPU emulation of the C-ABI entry points, written from the
semantics documented in include/r2n2_b200.h (channels-last buffers, packed correlation
weights, epilogue flag order, in-place conventions).

Purpose: (1) run the engine's host-side scheduling (tape, gradient accumulation, packing,
fusion flags) on the CPU box against the oracle, so that host logic is verified without a
GPU; (2) serve as a per-kernel expectation in the GPU tests.  Never imported by the product.
"""
import torch
import torch.nn.functional as F

LEAK = 0.01
EPI_BIAS, EPI_LRELU, EPI_RESIDUAL, EPI_MASK = 1, 2, 4, 8
EPI_RES_PRE = 64


def _f(t):
    return None if t is None else t.double()


CONV_UP2, CONV_DOWN2 = 16, 32


def _stuff(x5):
    """[N,D,H,W,C] -> zero-stuffed [N,2D,2H,2W,C] (value on the even corner, lib/layers.py:375-385)."""
    N, D, H, W, C = x5.shape
    o = torch.zeros(N, 2 * D, 2 * H, 2 * W, C, dtype=x5.dtype)
    o[:, ::2, ::2, ::2] = x5
    return o


def conv_fwd(x, w, bias, mask_src, residual, y, geom, cout, k, flags, algo, colsum=None):
    N, D, H, W, Cin = geom
    if flags & CONV_UP2:
        xi = _stuff(_f(x).view(N, D, H, W, Cin)).permute(0, 4, 1, 2, 3)
    elif flags & CONV_DOWN2:
        xi = _f(x).view(N, 2 * D, 2 * H, 2 * W, Cin).permute(0, 4, 1, 2, 3)
    else:
        xi = _f(x).view(N, D, H, W, Cin).permute(0, 4, 1, 2, 3)
    wi = _f(w).view(cout, k[0], k[1], k[2], Cin).permute(0, 4, 1, 2, 3)
    o = F.conv3d(xi, wi, None, padding=((k[0] - 1) // 2, (k[1] - 1) // 2, (k[2] - 1) // 2))
    if flags & CONV_DOWN2:
        o = o[:, :, ::2, ::2, ::2]
    o = o.permute(0, 2, 3, 4, 1).reshape(-1, cout)
    if (flags & EPI_RESIDUAL) and (flags & EPI_RES_PRE):
        o = o + _f(residual).view(-1, cout)
    if flags & EPI_BIAS:
        o = o + _f(bias).view(1, cout)
    if flags & EPI_LRELU:
        o = torch.where(o > 0, o, LEAK * o)
    if (flags & EPI_RESIDUAL) and not (flags & EPI_RES_PRE):
        o = o + _f(residual).view(-1, cout)
    if flags & EPI_MASK:
        m = _f(mask_src).view(-1, cout)
        o = o * torch.where(m > 0, torch.ones_like(m), torch.full_like(m, LEAK))
    y.copy_(o.reshape(-1).to(y.dtype))
    if colsum is not None:
        colsum.add_(y.view(-1, cout).double().sum(0).to(colsum.dtype))
    return y


def conv_wgrad(x, dy, dw, geom, cout, k, algo, accumulate, up2=False):
    N, D, H, W, Cin = geom
    kd, kh, kw = k
    pd, ph, pw = (kd - 1) // 2, (kh - 1) // 2, (kw - 1) // 2
    x5 = _f(x).view(N, D, H, W, Cin)
    if up2:
        x5 = _stuff(x5)
        D, H, W = 2 * D, 2 * H, 2 * W
    xi = F.pad(x5, (0, 0, pw, pw, ph, ph, pd, pd))
    g = _f(dy).view(N, D, H, W, cout)
    out = torch.zeros(cout, kd, kh, kw, Cin, dtype=torch.float64)
    for tz in range(kd):
        for ty in range(kh):
            for tx in range(kw):
                xs = xi[:, tz:tz + D, ty:ty + H, tx:tx + W]
                out[:, tz, ty, tx] = torch.einsum('ndhwo,ndhwc->oc', g, xs)
    out = out.reshape(-1).to(dw.dtype)
    if accumulate:
        dw.add_(out)
    else:
        dw.copy_(out)


def bias_grad(dy, db, C, accumulate):
    s = _f(dy).view(-1, C).sum(0).to(db.dtype)
    if accumulate:
        db.add_(s)
    else:
        db.copy_(s)


def weight_transpose(w, wt, cout, taps, cin):
    wt.copy_(w.view(cout, taps, cin).flip(1).permute(2, 1, 0).reshape(-1).to(wt.dtype))


def fill_zero(t):
    t.zero_()


def cast(src, dst):
    dst.copy_(src.to(dst.dtype))


def split_bf16(src, hi, lo):
    h = src.to(torch.bfloat16)
    hi.copy_(h)
    lo.copy_((src - h.float()).to(torch.bfloat16))


def nchw_to_nhwc(x, y, cpad, wpad=None, xoff=0):
    N, C, H, W = x.shape
    wpad = wpad or W
    o = torch.zeros(N, H, wpad, cpad, dtype=y.dtype)
    o[:, :, xoff:xoff + W, :C] = x.permute(0, 2, 3, 1).to(y.dtype)
    y.copy_(o.reshape(-1))
    return y


def nchw_to_rowpack(x, y, kw, cpack):
    N, C, H, W = x.shape
    pad = (kw - 1) // 2
    xp = F.pad(x, (pad, pad))
    o = torch.zeros(N, H, W, cpack, dtype=y.dtype)
    for kx in range(kw):
        o[..., kx * C:(kx + 1) * C] = xp[:, :, :, kx:kx + W].permute(0, 2, 3, 1).to(y.dtype)
    if kw * C < cpack:
        o[..., kw * C] = 1.0            # constant-one channel (bias gradient through the weight gradient)
    y.copy_(o.reshape(-1))
    return y


def _pool_in(a, b, N, H, W, C):
    s = _f(a).view(N, H, W, C)
    if b is not None:
        s = s + _f(b).view(N, H, W, C)
    return s


def maxpool_fwd(a, b, y, N, H, W, C, code=None):
    s = _pool_in(a, b, N, H, W, C).permute(0, 3, 1, 2)
    y.copy_(F.max_pool2d(s, 2, 2, 1).permute(0, 2, 3, 1).reshape(-1).to(y.dtype))
    if code is not None:        # r2n2_maxpool2d_fwd_code: bit k = window position k attains the maximum, bit 4+k = sign
        s = _pool_in(a, b, N, H, W, C)
        src = _f(a if b is None else b).view(N, H, W, C)
        Ho, Wo = H // 2 + 1, W // 2 + 1
        padding = (0, 0, 1, 2 * Wo - W - 1, 1, 2 * Ho - H - 1)
        win = F.pad(s, padding, value=float('-inf')).view(N, Ho, 2, Wo, 2, C)
        sg = F.pad(src, padding, value=0.0).view(N, Ho, 2, Wo, 2, C) > 0
        eq = win == win.amax(dim=(2, 4), keepdim=True)
        cb = torch.zeros(N, Ho, Wo, C, dtype=torch.int64)
        for dh in range(2):
            for dw in range(2):
                k = 2 * dh + dw
                cb |= eq[:, :, dh, :, dw].long() << k
                cb |= sg[:, :, dh, :, dw].long() << (4 + k)
        code.copy_(cb.reshape(-1).to(torch.uint8))


def maxpool_bwd_code(code, dy, dx, N, H, W, C, lrelu_mask, dxb=None, colsum_a=None, colsum_b=None):
    Ho, Wo = H // 2 + 1, W // 2 + 1
    cb = code.view(N, Ho, Wo, C).long()
    g = _f(dy).view(N, Ho, Wo, C)
    d = torch.zeros(N, Ho, 2, Wo, 2, C, dtype=g.dtype)
    sl = torch.zeros(N, Ho, 2, Wo, 2, C, dtype=g.dtype)
    for dh in range(2):
        for dw in range(2):
            k = 2 * dh + dw
            d[:, :, dh, :, dw] = torch.where(((cb >> k) & 1) == 1, g, torch.zeros_like(g))
            sl[:, :, dh, :, dw] = torch.where(((cb >> (4 + k)) & 1) == 1, torch.ones_like(g), torch.full_like(g, LEAK))
    d = d.reshape(N, 2 * Ho, 2 * Wo, C)[:, 1:1 + H, 1:1 + W]
    sl = sl.reshape(N, 2 * Ho, 2 * Wo, C)[:, 1:1 + H, 1:1 + W]
    dx.copy_((d * sl if lrelu_mask else d).reshape(-1).to(dx.dtype))
    if colsum_a is not None:
        colsum_a += _f(dx).view(-1, C).sum(0).to(colsum_a.dtype)
    if dxb is not None:
        dxb.copy_((d * sl).reshape(-1).to(dxb.dtype))
        if colsum_b is not None:
            colsum_b += _f(dxb).view(-1, C).sum(0).to(colsum_b.dtype)


def maxpool_bwd(a, b, dy, dx, N, H, W, C, lrelu_mask, dxb=None, colsum_a=None, colsum_b=None):
    s = _pool_in(a, b, N, H, W, C)
    Ho, Wo = H // 2 + 1, W // 2 + 1
    sp = F.pad(s, (0, 0, 1, 2 * Wo - W - 1, 1, 2 * Ho - H - 1), value=float('-inf'))
    win = sp.view(N, Ho, 2, Wo, 2, C)
    m = win.amax(dim=(2, 4), keepdim=True)
    g = _f(dy).view(N, Ho, 1, Wo, 1, C)
    d = torch.where(win == m, g.expand_as(win), torch.zeros_like(win)).reshape(N, 2 * Ho, 2 * Wo, C)
    d = d[:, 1:1 + H, 1:1 + W]
    if lrelu_mask:
        av = _f(a).view(N, H, W, C)
        d = d * torch.where(av > 0, torch.ones_like(av), torch.full_like(av, LEAK))
    dx.copy_(d.reshape(-1).to(dx.dtype))
    if colsum_a is not None:                        # sums of the values as stored
        colsum_a += _f(dx).view(-1, C).sum(0).to(colsum_a.dtype)
    if dxb is not None:
        bv = _f(b).view(N, H, W, C)
        db = d * torch.where(bv > 0, torch.ones_like(bv), torch.full_like(bv, LEAK))
        dxb.copy_(db.reshape(-1).to(dxb.dtype))
        if colsum_b is not None:
            colsum_b += _f(dxb).view(-1, C).sum(0).to(colsum_b.dtype)


def unpool_fwd(a, b, y, B, D, H, W, C):
    s = _f(a).view(B, D, H, W, C)
    if b is not None:
        s = s + _f(b).view(B, D, H, W, C)
    o = torch.zeros(B, 2 * D, 2 * H, 2 * W, C, dtype=torch.float64)
    o[:, ::2, ::2, ::2] = s
    y.copy_(o.reshape(-1).to(y.dtype))


def unpool_bwd(dy, dx, B, D, H, W, C):
    dx.copy_(dy.view(B, 2 * D, 2 * H, 2 * W, C)[:, ::2, ::2, ::2].reshape(-1))


def add(a, b, y):
    y.copy_((_f(a) + _f(b)).to(y.dtype))


def lrelu_fwd(x, y):
    v = _f(x)
    y.copy_(torch.where(v > 0, v, LEAK * v).to(y.dtype))


def lrelu_bwd(a, dy, addend, dx, colsum=None, C=None):
    av = _f(a)
    o = _f(dy) * torch.where(av > 0, torch.ones_like(av), torch.full_like(av, LEAK))
    if addend is not None:
        o = o + _f(addend)
    dx.copy_(o.to(dx.dtype))
    if colsum is not None:                          # sums of the values as stored
        colsum += _f(dx).view(-1, C).sum(0).to(colsum.dtype)


def eltwise(op, a, b, y):
    v = _f(a)
    if op < 3:
        o = [torch.sigmoid(v), torch.tanh(v), 1 - v][op]
    elif op == 3:
        o = v * _f(b)
    elif op == 4:
        o = _f(b) * v * (1 - v)
    elif op == 5:
        o = _f(b) * (1 - v * v)
    else:
        o = -v
    y.copy_(o.to(y.dtype))


def gru_ur_fwd(conv_ur, xp_ur, bias_ur, h, u, r, rh, B):
    a = _f(xp_ur).view(B * 64, 256) + _f(bias_ur).view(1, 256)
    hv = torch.zeros(B * 64, 128, dtype=torch.float64)
    if conv_ur is not None:
        a = a + _f(conv_ur).view(B * 64, 256)
        hv = _f(h).view(B * 64, 128)
    uu, rr = torch.sigmoid(a[:, :128]), torch.sigmoid(a[:, 128:])
    u.copy_(uu.reshape(-1).to(u.dtype)); r.copy_(rr.reshape(-1).to(r.dtype))
    rh.copy_((rr * hv).reshape(-1).to(rh.dtype))


def gru_out_fwd(conv_c, xp_c, bias_c, u, h, c, h_new, B):
    a = _f(xp_c).view(B * 64, 128) + _f(bias_c).view(1, 128)
    hv = torch.zeros(B * 64, 128, dtype=torch.float64)
    if conv_c is not None:
        a = a + _f(conv_c).view(B * 64, 128)
        hv = _f(h).view(B * 64, 128)
    cc = torch.tanh(a)
    uu = _f(u).view(B * 64, 128)
    c.copy_(cc.reshape(-1).to(c.dtype))
    h_new.copy_((uu * hv + (1 - uu) * cc).reshape(-1).to(h_new.dtype))


def gru_out_bwd(dh, u, c, h, da_ur, da_c, dh_prev, B):
    g, uu, cc = _f(dh).view(-1, 128), _f(u).view(-1, 128), _f(c).view(-1, 128)
    hv = torch.zeros_like(g) if h is None else _f(h).view(-1, 128)
    da_c.copy_((g * (1 - uu) * (1 - cc * cc)).reshape(-1).to(da_c.dtype))
    da_ur.view(-1, 256)[:, :128] = (g * (hv - cc) * uu * (1 - uu)).to(da_ur.dtype)
    dh_prev.copy_((g * uu).reshape(-1).to(dh_prev.dtype))


def gru_r_bwd(d_rh, h, r, da_ur, dh_prev, B):
    g, hv, rr = _f(d_rh).view(-1, 128), _f(h).view(-1, 128), _f(r).view(-1, 128)
    da_ur.view(-1, 256)[:, 128:] = (g * hv * rr * (1 - rr)).to(da_ur.dtype)
    dh_prev.copy_((_f(dh_prev).view(-1, 128) + g * rr).reshape(-1).to(dh_prev.dtype))


def conv_fwd_gru(mode, x, w, bias, ins, outs, geom, cout, k):
    """r2n2_conv_fwd_gru: hidden convolution (kept in float64, no rounding of the conv output) + gate math."""
    N, D, H, W, Cin = geom
    B = N * D * H * W // 64
    ins = list(ins) + [None] * (4 - len(ins))
    outs = list(outs) + [None] * (3 - len(outs))
    a = torch.empty(N * D * H * W * cout, dtype=torch.float64)
    conv_fwd(x, w, None, None, None, a, geom, cout, k, 0, 0)
    if mode == 1:
        gru_ur_fwd(a, ins[0], bias, ins[1], outs[0], outs[1], outs[2], B)
    elif mode == 2:
        gru_out_fwd(a, ins[0], bias, ins[1], ins[2], outs[0], outs[1], B)
    elif mode == 3:
        gru_r_bwd(a, ins[0], ins[1], outs[0], outs[1], B)
    else:
        dh = a + _f(ins[0])
        gru_out_bwd(dh, ins[1], ins[2], ins[3], outs[1], outs[0], outs[2], B)


def lstm_fwd(conv, xp, bias, s, fig, s_new, h_new, B):
    a = _f(xp).view(B * 64, 384) + _f(bias).view(1, 384)
    if conv is not None:
        a = a + _f(conv).view(B * 64, 384)
    sv = torch.zeros(B * 64, 128, dtype=torch.float64) if s is None else _f(s).view(B * 64, 128)
    f, i, g = torch.sigmoid(a[:, :128]), torch.sigmoid(a[:, 128:256]), torch.tanh(a[:, 256:])
    so = f * sv + i * g
    fig.copy_(torch.cat([f, i, g], 1).reshape(-1).to(fig.dtype))
    s_new.copy_(so.reshape(-1).to(s_new.dtype))
    h_new.copy_(torch.tanh(so).reshape(-1).to(h_new.dtype))


def lstm_bwd(dh, ds_in, fig, s, h_new, da, ds_prev, B):
    g, hv = _f(dh).view(-1, 128), _f(h_new).view(-1, 128)
    dsi = torch.zeros_like(g) if ds_in is None else _f(ds_in).view(-1, 128)
    sv = torch.zeros_like(g) if s is None else _f(s).view(-1, 128)
    fg = _f(fig).view(-1, 384)
    f, i, gg = fg[:, :128], fg[:, 128:256], fg[:, 256:]
    dst = g * (1 - hv * hv) + dsi
    out = torch.cat([dst * sv * f * (1 - f), dst * gg * i * (1 - i), dst * i * (1 - gg * gg)], 1)
    da.copy_(out.reshape(-1).to(da.dtype))
    ds_prev.copy_((dst * f).reshape(-1).to(ds_prev.dtype))


def softmax_ce3d(logits, y, pred, loss_sum, dlogits, B, nv, grad_scale, cstride=2):
    x = _f(logits).view(B, nv, nv, nv, cstride)[..., :2]
    e = torch.exp(x)
    p = e / e.sum(-1, keepdim=True)
    yy = torch.zeros_like(x) if y is None else _f(y).view(B, nv, 2, nv, nv).permute(0, 1, 3, 4, 2)
    if pred is not None:
        pred.copy_(p.permute(0, 1, 4, 2, 3).to(pred.dtype))
    if loss_sum is not None:
        loss_sum.add_((-(yy * x).sum(-1) + torch.log(e.sum(-1))).sum().to(loss_sum.dtype))
    if dlogits is not None:
        d = torch.zeros(B, nv, nv, nv, cstride, dtype=torch.float64)
        d[..., :2] = (p - yy) * grad_scale
        dlogits.copy_(d.reshape(-1).to(dlogits.dtype))


def adam(p, g, m, v, mirror, n_decay, lr_t, b1, b2, eps, wd, grad_scale):
    rg = g * grad_scale
    rg[:n_decay] += wd * p[:n_decay]
    m.mul_(b1).add_((1 - b1) * rg)
    v.mul_(b2).add_((1 - b2) * rg * rg)
    p.sub_(lr_t * m / (torch.sqrt(v) + eps))
    if mirror is not None:
        mirror.copy_(p.to(mirror.dtype))


def sgd(p, g, vel, mirror, n_decay, lr, momentum, wd, grad_scale):
    rg = g * grad_scale
    rg[:n_decay] += wd * p[:n_decay]
    vel.mul_(momentum).sub_(lr * rg)
    p.add_(vel)
    if mirror is not None:
        mirror.copy_(p.to(mirror.dtype))


def voxel_eval(pred, gt, thresh):
    occ = pred[:, :, 1] >= thresh
    g1, g0 = gt[:, :, 1] != 0, gt[:, :, 0] != 0
    f = lambda t: t.flatten(1).sum(1)
    return torch.stack([f(occ ^ g1), f(occ & g1), f(occ | g1), f(occ & g0), f(~occ & g1)], 1)


ALL = ['conv_fwd_gru', 'conv_fwd', 'conv_wgrad', 'bias_grad', 'weight_transpose', 'cast', 'split_bf16', 'fill_zero', 'nchw_to_nhwc', 'nchw_to_rowpack',
       'maxpool_fwd', 'maxpool_bwd', 'maxpool_bwd_code', 'unpool_fwd', 'unpool_bwd', 'add', 'lrelu_fwd', 'lrelu_bwd',
       'eltwise', 'gru_ur_fwd', 'gru_out_fwd', 'gru_out_bwd', 'gru_r_bwd', 'lstm_fwd', 'lstm_bwd',
       'softmax_ce3d', 'adam', 'sgd', 'voxel_eval']


def install(monkeypatch, ops_module):
    """Route r2n2_b200.ops raw wrappers to the emulation (CPU host-logic tests only)."""
    import sys
    me = sys.modules[__name__]
    for name in ALL:
        monkeypatch.setattr(ops_module, name, getattr(me, name))

"""Reference-layout <-> packed-layout transforms for parameters and activations (pure torch;
runs on CPU or GPU; only used at load/save/set_value time, never in the timed path).

Reference layouts (lib/layers.py): conv2d W (Cout,Cin,kh,kw) [282-284]; conv3d W / Wh
(Cout,kd,Cin,kh,kw) [394-398, 452-456]; fc W (n_in,n_out) with n_in flattened (C,H,W)-major
[140, 235-236]; Wx (n_fc, D*Cout*H*W) with columns in (d,c,h,w) order [469, 502-503].
Theano's conv2d / conv3d2d.conv3d are TRUE convolutions (kernel flipped on every spatial
axis); the kernels in csrc/ compute correlations over channels-last tensors, so packing
flips the taps once and moves the channel axis last.
"""
import torch


def pack_conv2d(w_ref, cin_pad=None):
    """(Cout,Cin,kh,kw) -> [Cout,kh,kw,Cin_pad]"""
    cout, cin, kh, kw = w_ref.shape
    p = w_ref.flip(2, 3).permute(0, 2, 3, 1)
    cin_pad = cin_pad or cin
    if cin_pad != cin:
        out = w_ref.new_zeros(cout, kh, kw, cin_pad)
        out[..., :cin] = p
        return out
    return p.contiguous()


def unpack_conv2d(w_packed, cin):
    """[Cout,kh,kw,Cin_pad] -> (Cout,Cin,kh,kw)"""
    return w_packed[..., :cin].permute(0, 3, 1, 2).flip(2, 3).contiguous()


def pack_conv2d_rowpack(w_ref, cpack):
    """(Cout,Cin,kh,kw) -> [Cout,kh,cpack] for a row-packed input (channel j = kx*Cin + c holds
    x[..., w + kx - pad]): packed[co,ty,kx*Cin+c] = w_ref[co,c,kh-1-ty,kw-1-kx], zeros for j >= kw*Cin."""
    cout, cin, kh, kw = w_ref.shape
    p = w_ref.flip(2, 3).permute(0, 2, 3, 1).reshape(cout, kh, kw * cin)     # [co,ty,kx*cin+c]
    out = w_ref.new_zeros(cout, kh, cpack)
    out[..., :kw * cin] = p
    return out


def unpack_conv2d_rowpack(w_packed, cin, kw):
    """[Cout,kh,cpack] -> (Cout,Cin,kh,kw)"""
    cout, kh, _ = w_packed.shape
    p = w_packed[..., :kw * cin].reshape(cout, kh, kw, cin)
    return p.permute(0, 3, 1, 2).flip(2, 3).contiguous()


def pack_conv3d(w_ref):
    """(Cout,kd,Cin,kh,kw) -> [Cout,kd,kh,kw,Cin]"""
    return w_ref.flip(1, 3, 4).permute(0, 1, 3, 4, 2).contiguous()


def unpack_conv3d(w_packed):
    """[Cout,kd,kh,kw,Cin] -> (Cout,kd,Cin,kh,kw)"""
    return w_packed.permute(0, 1, 4, 2, 3).flip(1, 3, 4).contiguous()


def pack_fc(w_ref, chw=None):
    """(n_in,n_out) -> [n_out, n_in'] where n_in' is (H,W,C)-ordered if the input was a flattened
    (C,H,W) feature map (chw=(C,H,W)), else unchanged."""
    n_in, n_out = w_ref.shape
    if chw is None:
        return w_ref.t().contiguous()
    c, h, w = chw
    assert c * h * w == n_in
    return w_ref.reshape(c, h, w, n_out).permute(3, 1, 2, 0).reshape(n_out, n_in).contiguous()


def unpack_fc(w_packed, chw=None):
    n_out, n_in = w_packed.shape
    if chw is None:
        return w_packed.t().contiguous()
    c, h, w = chw
    return w_packed.reshape(n_out, h, w, c).permute(3, 1, 2, 0).reshape(n_in, n_out).contiguous()


def pack_fcconv_wx(wx_ref, d, c, h, w):
    """(n_fc, d*c*h*w) (d,c,h,w)-ordered columns -> [d*h*w, c, n_fc]  (rows = GEMM N index
    pos*c + ch in channels-last hidden-grid order, K = n_fc contiguous)."""
    n_fc = wx_ref.shape[0]
    return wx_ref.reshape(n_fc, d, c, h, w).permute(1, 3, 4, 2, 0).reshape(d * h * w, c, n_fc)


def unpack_fcconv_wx(wx_packed, d, c, h, w):
    """[d*h*w, c, n_fc] (possibly a strided view) -> (n_fc, d*c*h*w)"""
    n_fc = wx_packed.shape[-1]
    return wx_packed.unflatten(0, (d, h, w)).permute(4, 0, 3, 1, 2).reshape(n_fc, d * c * h * w).contiguous()


# ---- activations -----------------------------------------------------------------------
def to_channels_last_2d(x_ref):
    """(N,C,H,W) -> [N,H,W,C] contiguous"""
    return x_ref.permute(0, 2, 3, 1).contiguous()


def ref_view_2d(x_cl):
    """[N,H,W,C] -> logical (N,C,H,W) view (no copy)"""
    return x_cl.permute(0, 3, 1, 2)


def to_channels_last_3d(x_ref):
    """(B,D,C,H,W) -> [B,D,H,W,C] contiguous"""
    return x_ref.permute(0, 1, 3, 4, 2).contiguous()


def ref_view_3d(x_cl):
    """[B,D,H,W,C] -> logical (B,D,C,H,W) view (no copy)"""
    return x_cl.permute(0, 1, 4, 2, 3)

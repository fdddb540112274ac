"""Host-side op layer: thin wrappers that turn torch CUDA tensors into C-ABI calls
(include/r2n2_b200.h) plus a minimal reverse-mode tape so that the reference's
`tensor.grad(loss, params)` (models/net.py:58) has an equivalent.

torch is used for device memory and streams only: every FLOP and every byte moved on the
hot path goes through libr2n2_b200.so.  There is no CPU path here: tensors must be CUDA
tensors and the library must load.
"""
import os

import torch

from . import _lib as L

_DT = {torch.float32: L.F32, torch.bfloat16: L.BF16}

launch_count = 0  # number of r2n2_* kernels launched by this process (bench.py reports it)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


class Profiler:
    """Optional per-launch CUDA-event timing of every C-ABI call (bench.py's roofline pass).
    Events are recorded on the launching (current) stream; resolve() after a synchronize."""

    def __init__(self):
        self.records = []       # (name, tag, work, start_event, end_event)

    def resolve(self):
        """-> {name: dict(ms, calls, flops, bytes)} plus the raw per-launch list."""
        table, raw = {}, []
        for name, tag, work, e0, e1 in self.records:
            ms = e0.elapsed_time(e1)
            d = table.setdefault(name, dict(ms=0.0, calls=0, flops=0.0, bytes=0.0))
            d['ms'] += ms
            d['calls'] += 1
            d['flops'] += work.get('flops', 0.0)
            d['bytes'] += work.get('bytes', 0.0)
            raw.append((name, tag, ms, work))
        return table, raw


profiler = None          # set to a Profiler() to time every launch
_next_work = None        # work description (flops / bytes) attached by the wrapper about to launch
_next_tag = ''


pre_launch = None         # optional host callback before every launch (dist.GradSync polls its all-reduces here)


def _call(name, *args):
    global launch_count, _next_work, _next_tag
    launch_count += 1
    if pre_launch is not None:
        pre_launch()
    if profiler is None:
        L.call(name, *args)
        return
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    L.call(name, *args)
    e1.record()
    profiler.records.append((name, _next_tag, _next_work or {}, e0, e1))
    _next_work, _next_tag = None, ''


def _work(flops=0.0, tensors=()):
    """Describe the launch that follows: algorithmic flops and bytes (sum of operand sizes)."""
    global _next_work
    if profiler is not None:
        nbytes = 0
        for t in tensors:
            if t is not None:
                nbytes += t.numel() * t.element_size()
        _next_work = dict(flops=float(flops), bytes=float(nbytes))


WGRAD_UMMA = True        # tcgen05 weight-gradient kernel available
FUSE_BIAS_GRAD = True     # bias gradients as column sums of the data-gradient epilogue that produces dy
TRANSPOSE_FROM_MIRROR = True   # data-gradient operands are transposed from the bf16 mirror (half the reads)
FUSE_POOL_BWD = True      # max-pool backward also writes the rectified branch's pre-activation gradient + bias sums
FUSE_LRELU_BIAS = not os.environ.get('R2N2_NO_LRELU_BIAS')    # standalone LeakyReLU' passes also emit the bias column sums
POOL_CODES = not os.environ.get('R2N2_NO_POOL_CODES')   # training: max-pool backward from arg-max/sign bytes (no re-read of a, b)
ZERO_SKIP_ANY_DTYPE = False   # host-logic tests on the emulated ABI only (the kernels are bf16)
useful_frac = 1.0        # fraction of the dense MACs of the next conv that are algorithmically
                         # useful (1/8 on zero-stuffed Unpool3D inputs, 3/Cpad on the padded RGB)


def _check(t, dtype=None):
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError('r2n2_b200 ops need CUDA tensors (there is no CPU fallback)')
    if not t.is_contiguous():
        raise RuntimeError('r2n2_b200 ops need contiguous tensors, got strides %s' % (t.stride(),))
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError('dtype mismatch: expected %s got %s' % (dtype, t.dtype))


def umma_ok(cin, cout):
    """Shapes the tcgen05 kernel takes (16-element K and N granules); anything else (the RGB-free
    conv11 with Cout=2 and its data gradient with Cin=2) runs on the fp32-FFMA kernel."""
    return cin % 16 == 0 and cout % 16 == 0


def umma_error_flag():
    return L.load().r2n2_umma_error_flag()


# ======================================================================= raw op wrappers
def conv_fwd(x, w, bias, mask_src, residual, y, geom, cout, k, flags, algo, colsum=None):
    """geom = (N,D,H,W,Cin) of x; k=(kd,kh,kw); y preallocated [N,D,H,W,cout].
    With L.CONV_UP2 / L.CONV_DOWN2 in flags (N,D,H,W) is the SMALL grid: UP2 writes y (and reads
    mask/residual) at [N,2D,2H,2W,cout]; DOWN2 reads x at [N,2D,2H,2W,Cin].
    colsum (fp32 [cout]): additionally colsum[c] += sum_p y[p,c] (bias gradient of the producing layer)."""
    N, D, H, W, Cin = geom
    if algo == L.ALGO_UMMA and not (umma_ok(Cin, cout) and x.dtype == torch.bfloat16):
        algo = L.ALGO_SIMT
    for t in (x, w, bias, mask_src, residual, y):
        _check(t)
    xs = 8 if flags & L.CONV_DOWN2 else 1
    ys = 8 if flags & L.CONV_UP2 else 1
    assert x.numel() == N * D * H * W * Cin * xs, (x.shape, geom)
    assert w.numel() == cout * k[0] * k[1] * k[2] * Cin, (w.shape, cout, k, Cin)
    assert y.numel() == N * D * H * W * cout * ys
    assert w.dtype == x.dtype
    global _next_tag
    _work(2.0 * N * D * H * W * Cin * cout * k[0] * k[1] * k[2] * useful_frac,
          (x, w, bias, mask_src, residual, y))
    _next_tag = 'M%dxN%dxK%d' % (N * D * H * W, cout, k[0] * k[1] * k[2] * Cin)
    if colsum is not None:
        _check(colsum, torch.float32)
        assert colsum.numel() == cout
        _call('r2n2_conv_fwd_colsum', _ptr(x), _ptr(w), _ptr(bias), _ptr(mask_src), _ptr(residual), _ptr(y),
              N, D, H, W, Cin, cout, k[0], k[1], k[2], flags, _DT[x.dtype], _DT[y.dtype], algo,
              _ptr(colsum), _stream())
        return y
    _call('r2n2_conv_fwd', _ptr(x), _ptr(w), _ptr(bias), _ptr(mask_src), _ptr(residual), _ptr(y),
          N, D, H, W, Cin, cout, k[0], k[1], k[2], flags, _DT[x.dtype], _DT[y.dtype], algo,
          _stream())
    return y


def conv_wgrad(x, dy, dw, geom, cout, k, algo, accumulate, up2=False):
    """up2: x is the small [N,D,H,W,Cin] input of an UP2 convolution, dy is [N,2D,2H,2W,cout]."""
    N, D, H, W, Cin = geom
    if algo == L.ALGO_UMMA and not (umma_ok(Cin, cout) and x.dtype == torch.bfloat16 and WGRAD_UMMA):
        algo = L.ALGO_SIMT
    for t in (x, dy, dw):
        _check(t)
    assert dw.dtype == torch.float32 and x.dtype == dy.dtype
    assert x.numel() == N * D * H * W * Cin and dy.numel() == N * D * H * W * cout * (8 if up2 else 1)
    assert dw.numel() == cout * k[0] * k[1] * k[2] * Cin
    global _next_tag
    _work(2.0 * N * D * H * W * Cin * cout * k[0] * k[1] * k[2] * useful_frac, (x, dy, dw))
    _next_tag = 'M%dxN%dxK%d' % (cout, k[0] * k[1] * k[2] * Cin, N * D * H * W)
    _call('r2n2_conv_wgrad', _ptr(x), _ptr(dy), _ptr(dw), N, D, H, W, Cin, cout, k[0], k[1], k[2],
          _DT[x.dtype], algo, (L.WGRAD_ACCUMULATE if accumulate else 0) | (L.WGRAD_UP2 if up2 else 0),
          _stream())


# ======================================================================= hi/lo operand split (COMPUTE='bf16x3')
class SplitPair:
    """(hi, lo) bf16 tensors with hi + lo == the fp32 source up to 2^-17 relative."""
    __slots__ = ('hi', 'lo', 'src')

    def __init__(self, hi, lo, src=None):
        self.hi, self.lo, self.src = hi, lo, src


def split_bf16(src, hi, lo):
    _check(src, torch.float32), _check(hi, torch.bfloat16), _check(lo, torch.bfloat16)
    assert hi.numel() == src.numel() == lo.numel()
    _work(0.0, (src, hi, lo))
    _call('r2n2_split_bf16', _ptr(src), _ptr(hi), _ptr(lo), src.numel(), _stream())


# activations split during the current step: (data_ptr, numel) -> SplitPair.  The pair keeps a reference to its
# source, so the allocator cannot hand the same address to another tensor while the entry lives; cleared by the
# engine at the start of every forward pass and at the end of every backward pass.
_split_cache = {}


def split_cache_clear():
    _split_cache.clear()


def split_invalidate(t):
    """t was overwritten in place (gradient accumulation, LeakyReLU' applied to a handed-over buffer): a split
    taken before the write is stale."""
    if _split_cache and t is not None:
        _split_cache.pop((t.data_ptr(), t.numel()), None)


def split_of(x):
    """x: fp32 tensor (an activation / gradient that is complete) or an existing SplitPair."""
    if isinstance(x, SplitPair):
        return x
    key = (x.data_ptr(), x.numel())
    sp = _split_cache.get(key)
    if sp is None:
        hi = torch.empty(x.numel(), dtype=torch.bfloat16, device=x.device)
        lo = torch.empty(x.numel(), dtype=torch.bfloat16, device=x.device)
        split_bf16(x, hi, lo)
        sp = SplitPair(hi, lo, x)
        _split_cache[key] = sp
    return sp


def conv_fwd_x(x, w, bias, mask_src, residual, y, geom, cout, k, flags, algo, colsum=None):
    """conv_fwd, or -- algo == L.ALGO_SPLIT -- the same contraction at fp32 accuracy on the tensor cores:
    x and w are split into bf16 (hi, lo) pairs and  hi*hi + lo*hi + hi*lo  is accumulated over three tcgen05
    passes into the fp32 output (R2N2_EPI_RES_PRE: the partial sum enters before bias / LeakyReLU / mask;
    the dropped lo*lo term is 2^-18 relative).  w is then a SplitPair of packed operands."""
    if algo != L.ALGO_SPLIT:
        return conv_fwd(x, w, bias, mask_src, residual, y, geom, cout, k, flags, algo, colsum)
    Cin = geom[4]
    assert isinstance(w, SplitPair) and y.dtype == torch.float32
    if not umma_ok(Cin, cout):
        raise RuntimeError('bf16x3: Cin %d / Cout %d is not a tcgen05 shape (multiples of 16)' % (Cin, cout))
    xs = split_of(x)
    global useful_frac
    uf = useful_frac
    useful_frac = uf / 3.0          # (profiler bookkeeping: the three passes are ONE algorithmic contraction)
    try:
        return _conv_fwd_split(xs, w, bias, mask_src, residual, y, geom, cout, k, flags, colsum)
    finally:
        useful_frac = uf


def _conv_fwd_split(xs, w, bias, mask_src, residual, y, geom, cout, k, flags, colsum):
    geo = flags & (L.CONV_UP2 | L.CONV_DOWN2)
    real_res = residual if flags & L.EPI_RESIDUAL else None
    assert not (real_res is not None and flags & L.EPI_LRELU), 'a post-activation residual cannot ride the split passes'
    conv_fwd(xs.hi, w.hi, None, None, real_res, y, geom, cout, k, geo | (L.EPI_RESIDUAL if real_res is not None else 0),
             L.ALGO_UMMA)
    conv_fwd(xs.lo, w.hi, None, None, y, y, geom, cout, k, geo | L.EPI_RESIDUAL, L.ALGO_UMMA)
    conv_fwd(xs.hi, w.lo, bias, mask_src, y, y, geom, cout, k,
             geo | L.EPI_RESIDUAL | L.EPI_RES_PRE | (flags & (L.EPI_BIAS | L.EPI_LRELU | L.EPI_MASK)), L.ALGO_UMMA,
             colsum)
    return y


def conv_wgrad_x(x, dy, dw, geom, cout, k, algo, accumulate, up2=False):
    if algo != L.ALGO_SPLIT:
        return conv_wgrad(x, dy, dw, geom, cout, k, algo, accumulate, up2)
    if not umma_ok(geom[4], cout):
        raise RuntimeError('bf16x3: Cin %d / Cout %d is not a tcgen05 shape (multiples of 16)' % (geom[4], cout))
    xs, ds = split_of(x), split_of(dy)
    global useful_frac
    uf = useful_frac
    useful_frac = uf / 3.0
    try:
        conv_wgrad(xs.hi, ds.hi, dw, geom, cout, k, L.ALGO_UMMA, accumulate, up2)
        conv_wgrad(xs.lo, ds.hi, dw, geom, cout, k, L.ALGO_UMMA, True, up2)
        conv_wgrad(xs.hi, ds.lo, dw, geom, cout, k, L.ALGO_UMMA, True, up2)
    finally:
        useful_frac = uf


def bias_grad(dy, db, C, accumulate):
    _check(dy), _check(db, torch.float32)
    _work(0.0, (dy,))
    _call('r2n2_bias_grad', _ptr(dy), _ptr(db), dy.numel() // C, C, _DT[dy.dtype],
          1 if accumulate else 0, _stream())


def weight_transpose(w, wt, cout, taps, cin):
    _check(w, torch.float32), _check(wt)
    _work(0.0, (w, wt))
    _call('r2n2_weight_transpose', _ptr(w), _ptr(wt), cout, taps, cin, _DT[wt.dtype], _stream())


class TransposeBatch:
    """Transposed (data-gradient) operands of ALL weight tensors of a flat store, refreshed by one
    launch per parameter version instead of one launch per layer."""

    def __init__(self, params, store, mirror_attr='mirror', wt_attr='wt'):
        self.store = store
        self.params = params
        self.mirror_attr = mirror_attr            # 'mirror' (bf16 / hi operands) or 'mirror_lo' (bf16x3 lo operands)
        mirror = getattr(store, mirror_attr)
        dt = mirror.dtype if mirror is not None else torch.float32
        rows, dst, tiles = [], 0, 0
        for p in params:
            rows.append([p.data.storage_offset(), dst, p.cout, p.taps, p.cin, tiles])
            dst += (p.data.numel() + 63) // 64 * 64
            tiles += ((p.cin + 31) // 32) * ((p.cout + 63) // 64) * p.taps
        self.total_tiles = tiles
        self.table = torch.tensor(rows, dtype=torch.int64, device=store.p.device)
        # from the bf16 mirror (same offsets): 64 x 64 tiles, needs even offsets / channel counts
        self.table64 = None
        if (dt == torch.bfloat16 and TRANSPOSE_FROM_MIRROR
                and all(r[0] % 2 == 0 and r[1] % 2 == 0 and r[2] % 2 == 0 and r[4] % 2 == 0 for r in rows)):
            rows64, tiles64 = [], 0
            for r in rows:
                rows64.append(r[:5] + [tiles64])
                tiles64 += ((r[4] + 63) // 64) * ((r[2] + 63) // 64) * r[3]
            self.total_tiles64 = tiles64
            self.table64 = torch.tensor(rows64, dtype=torch.int64, device=store.p.device)
        self.wt_all = torch.empty(dst, dtype=dt, device=store.p.device)
        for p, r in zip(params, rows):
            setattr(p, wt_attr, self.wt_all[r[1]:r[1] + p.data.numel()])
            if wt_attr == 'wt':
                p.tbatch = self
            else:
                p.tbatch_lo = self
        self.version = -1

    def ensure(self):
        if self.version != self.store.version:
            if self.table64 is not None:
                mirror = getattr(self.store, self.mirror_attr)
                _check(mirror, torch.bfloat16)
                _work(0.0, (mirror, self.wt_all))
                _call('r2n2_weight_transpose_batch_bf16', _ptr(mirror), _ptr(self.wt_all),
                      _ptr(self.table64), len(self.params), self.total_tiles64, _stream())
            else:
                _check(self.store.p, torch.float32)
                _work(0.0, (self.store.p, self.wt_all))
                _call('r2n2_weight_transpose_batch', _ptr(self.store.p), _ptr(self.wt_all), _ptr(self.table),
                      len(self.params), self.total_tiles, _DT[self.wt_all.dtype], _stream())
            self.version = self.store.version


def fill_zero(t):
    _check(t)
    _call('r2n2_fill_zero', _ptr(t), t.numel() * t.element_size(), _stream())


def cast(src, dst):
    _check(src, torch.float32), _check(dst)
    _work(0.0, (src, dst))
    _call('r2n2_cast', _ptr(src), _ptr(dst), src.numel(), _DT[dst.dtype], _stream())


def nchw_to_nhwc(x, y, cpad, wpad=None, xoff=0):
    N, C, H, W = x.shape
    _check(x, torch.float32), _check(y)
    wpad = wpad or W
    assert y.numel() == N * H * wpad * cpad
    _call('r2n2_nchw_to_nhwc', _ptr(x), _ptr(y), N, C, H, W, cpad, wpad, xoff, _DT[y.dtype],
          _stream())
    return y


def nchw_to_rowpack(x, y, kw, cpack):
    N, C, H, W = x.shape
    _check(x, torch.float32), _check(y)
    assert y.numel() == N * H * W * cpack
    _call('r2n2_nchw_to_rowpack', _ptr(x), _ptr(y), N, C, H, W, kw, cpack, _DT[y.dtype], _stream())
    return y


def maxpool_fwd(a, b, y, N, H, W, C, code=None):
    """code (uint8 [N,Ho,Wo,C], optional): arg-max / sign byte per output element for maxpool_bwd_code."""
    _check(a), _check(b), _check(y)
    if code is not None:
        _check(code, torch.uint8)
        assert code.numel() == y.numel()
        _work(0.0, (a, b, y, code))
        _call('r2n2_maxpool2d_fwd_code', _ptr(a), _ptr(b), _ptr(y), _ptr(code), N, H, W, C, _DT[a.dtype], _stream())
        return
    _work(0.0, (a, b, y))
    _call('r2n2_maxpool2d_fwd', _ptr(a), _ptr(b), _ptr(y), N, H, W, C, _DT[a.dtype], _stream())


def maxpool_bwd_code(code, dy, dx, N, H, W, C, lrelu_mask, dxb=None, colsum_a=None, colsum_b=None):
    """maxpool_bwd from the forward pass's code bytes: neither a nor b is read again."""
    _check(code, torch.uint8)
    for t in (dy, dx, dxb):
        _check(t)
    for t in (colsum_a, colsum_b):
        if t is not None:
            _check(t, torch.float32)
            assert t.numel() == C
    assert code.numel() == dy.numel()
    _work(0.0, (code, dy, dx, dxb))
    _call('r2n2_maxpool2d_bwd_code', _ptr(code), _ptr(dy), _ptr(dx), _ptr(dxb), _ptr(colsum_a), _ptr(colsum_b),
          N, H, W, C, 1 if lrelu_mask else 0, _DT[dy.dtype], _stream())


def maxpool_bwd(a, b, dy, dx, N, H, W, C, lrelu_mask, dxb=None, colsum_a=None, colsum_b=None):
    """dxb (optional, needs b): dx * LeakyReLU'(b) written by the same kernel; colsum_a / colsum_b (fp32 [C]):
    += column sums of dx / dxb (the bias gradients of the layers that produced a and b)."""
    for t in (a, b, dy, dx, dxb):
        _check(t)
    for t in (colsum_a, colsum_b):
        if t is not None:
            _check(t, torch.float32)
            assert t.numel() == C
    _work(0.0, (a, b, dy, dx, dxb))
    if dxb is None and colsum_a is None and colsum_b is None:
        _call('r2n2_maxpool2d_bwd', _ptr(a), _ptr(b), _ptr(dy), _ptr(dx), N, H, W, C,
              1 if lrelu_mask else 0, _DT[a.dtype], _stream())
    else:
        _call('r2n2_maxpool2d_bwd_fused', _ptr(a), _ptr(b), _ptr(dy), _ptr(dx), _ptr(dxb), _ptr(colsum_a),
              _ptr(colsum_b), N, H, W, C, 1 if lrelu_mask else 0, _DT[a.dtype], _stream())


def unpool_fwd(a, b, y, B, D, H, W, C):
    _check(a), _check(b), _check(y)
    _work(0.0, (a, b, y))
    _call('r2n2_unpool3d_fwd', _ptr(a), _ptr(b), _ptr(y), B, D, H, W, C, _DT[a.dtype], _stream())


def unpool_bwd(dy, dx, B, D, H, W, C):
    _check(dy), _check(dx)
    _work(0.0, (dy, dx))
    _call('r2n2_unpool3d_bwd', _ptr(dy), _ptr(dx), B, D, H, W, C, _DT[dy.dtype], _stream())


def add(a, b, y):
    _check(a), _check(b), _check(y)
    _work(0.0, (a, b, y))
    _call('r2n2_add', _ptr(a), _ptr(b), _ptr(y), a.numel(), _DT[a.dtype], _stream())


def lrelu_fwd(x, y):
    _check(x), _check(y)
    _call('r2n2_lrelu_fwd', _ptr(x), _ptr(y), x.numel(), _DT[x.dtype], _stream())


def lrelu_bwd(a, dy, addend, dx, colsum=None, C=None):
    """colsum (fp32 [C], optional; a is then [rows, C]): += column sums of dx (bias gradient of the conv that produced a)."""
    for t in (a, dy, addend, dx):
        _check(t)
    _work(0.0, (a, dy, addend, dx))
    if colsum is not None:
        _check(colsum, torch.float32)
        assert colsum.numel() == C and a.numel() % C == 0
        _call('r2n2_lrelu_bwd_colsum', _ptr(a), _ptr(dy), _ptr(addend), _ptr(dx), a.numel() // C, C, _ptr(colsum),
              _DT[a.dtype], _stream())
        return
    _call('r2n2_lrelu_bwd', _ptr(a), _ptr(dy), _ptr(addend), _ptr(dx), a.numel(), _DT[a.dtype],
          _stream())


def eltwise(op, a, b, y):
    _check(a), _check(b), _check(y)
    _call('r2n2_eltwise', op, _ptr(a), _ptr(b), _ptr(y), a.numel(), _DT[a.dtype], _stream())


def conv_fwd_gru(mode, x, w, bias, ins, outs, geom, cout, k):
    """Hidden-state convolution of the conv-GRU with the gate math in its epilogue (r2n2_conv_fwd_gru, tcgen05 path, bf16):
    mode = L.GRU_UR / GRU_BLEND / GRU_RBWD / GRU_OBWD; ins (<= 4) / outs (<= 3) as documented in include/r2n2_b200.h."""
    N, D, H, W, Cin = geom
    ins = list(ins) + [None] * (4 - len(ins))
    outs = list(outs) + [None] * (3 - len(outs))
    for t in [x, w] + ins + outs:
        _check(t, torch.bfloat16)
    _check(bias, torch.float32)
    assert x.numel() == N * D * H * W * Cin and w.numel() == cout * k[0] * k[1] * k[2] * Cin
    global _next_tag
    _work(2.0 * N * D * H * W * Cin * cout * k[0] * k[1] * k[2], [x, w] + ins + outs)
    _next_tag = 'M%dxN%dxK%d' % (N * D * H * W, cout, k[0] * k[1] * k[2] * Cin)
    _call('r2n2_conv_fwd_gru', mode, _ptr(x), _ptr(w), _ptr(bias), *[_ptr(t) for t in ins], *[_ptr(t) for t in outs],
          N, D, H, W, Cin, cout, k[0], k[1], k[2], L.BF16, _stream())


def gru_ur_fwd(conv_ur, xp_ur, bias_ur, h, u, r, rh, B):
    _work(0.0, (conv_ur, xp_ur, h, u, r, rh))
    _call('r2n2_gru_gates_ur_fwd', _ptr(conv_ur), _ptr(xp_ur), _ptr(bias_ur), _ptr(h), _ptr(u),
          _ptr(r), _ptr(rh), B, _DT[xp_ur.dtype], _stream())


def gru_out_fwd(conv_c, xp_c, bias_c, u, h, c, h_new, B):
    _work(0.0, (conv_c, xp_c, u, h, c, h_new))
    _call('r2n2_gru_gates_out_fwd', _ptr(conv_c), _ptr(xp_c), _ptr(bias_c), _ptr(u), _ptr(h),
          _ptr(c), _ptr(h_new), B, _DT[xp_c.dtype], _stream())


def gru_out_bwd(dh, u, c, h, da_ur, da_c, dh_prev, B):
    _work(0.0, (dh, u, c, h, da_c, dh_prev, dh_prev))
    _call('r2n2_gru_gates_out_bwd', _ptr(dh), _ptr(u), _ptr(c), _ptr(h), _ptr(da_ur), _ptr(da_c),
          _ptr(dh_prev), B, _DT[dh.dtype], _stream())


def gru_r_bwd(d_rh, h, r, da_ur, dh_prev, B):
    _work(0.0, (d_rh, h, r, dh_prev, dh_prev, d_rh))
    _call('r2n2_gru_gates_r_bwd', _ptr(d_rh), _ptr(h), _ptr(r), _ptr(da_ur), _ptr(dh_prev), B,
          _DT[d_rh.dtype], _stream())


def lstm_fwd(conv, xp, bias, s, fig, s_new, h_new, B):
    _work(0.0, (conv, xp, s, fig, s_new, h_new))
    _call('r2n2_lstm_gates_fwd', _ptr(conv), _ptr(xp), _ptr(bias), _ptr(s), _ptr(fig), _ptr(s_new),
          _ptr(h_new), B, _DT[xp.dtype], _stream())


def lstm_bwd(dh, ds_in, fig, s, h_new, da, ds_prev, B):
    _work(0.0, (dh, ds_in, fig, s, h_new, da, ds_prev))
    _call('r2n2_lstm_gates_bwd', _ptr(dh), _ptr(ds_in), _ptr(fig), _ptr(s), _ptr(h_new), _ptr(da),
          _ptr(ds_prev), B, _DT[dh.dtype], _stream())


def softmax_ce3d(logits, y, pred, loss_sum, dlogits, B, nv, grad_scale, cstride=2):
    _check(logits, torch.float32), _check(y, torch.float32), _check(pred, torch.float32)
    _check(loss_sum, torch.float32), _check(dlogits)
    dd = _DT[dlogits.dtype] if dlogits is not None else L.F32
    _work(0.0, (logits, y, pred, dlogits))
    _call('r2n2_softmax_ce3d', _ptr(logits), _ptr(y), _ptr(pred), _ptr(loss_sum), _ptr(dlogits), B,
          nv, float(grad_scale), dd, cstride, _stream())


def adam(p, g, m, v, mirror, n_decay, lr_t, b1, b2, eps, wd, grad_scale):
    for t in (p, g, m, v):
        _check(t, torch.float32)
    _check(mirror)
    _work(0.0, (p, g, m, v, p, m, v, mirror))
    _call('r2n2_adam', _ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(mirror), p.numel(), n_decay,
          float(lr_t), float(b1), float(b2), float(eps), float(wd), float(grad_scale),
          _DT[mirror.dtype] if mirror is not None else L.F32, _stream())


def sgd(p, g, vel, mirror, n_decay, lr, momentum, wd, grad_scale):
    for t in (p, g, vel):
        _check(t, torch.float32)
    _call('r2n2_sgd', _ptr(p), _ptr(g), _ptr(vel), _ptr(mirror), p.numel(), n_decay, float(lr),
          float(momentum), float(wd), float(grad_scale),
          _DT[mirror.dtype] if mirror is not None else L.F32, _stream())


def voxel_eval(pred, gt, thresh):
    """lib/voxel.py:4-11 on the GPU: returns int64 tensor [B,5]."""
    _check(pred, torch.float32), _check(gt, torch.float32)
    B, nv = pred.shape[0], pred.shape[1]
    counts = torch.empty(B, 5, dtype=torch.int64, device=pred.device)
    _call('r2n2_voxel_eval', _ptr(pred), _ptr(gt), float(thresh), _ptr(counts), B, nv, _stream())
    return counts


def preprocess_views(rgba, params, out):
    """rgba uint8 [n,Hs,Ws,4], params int32 [n,6], out [n,3,Ho,Wo] (fp32 / bf16): lib/data_augmentation.py:57-68."""
    _check(rgba, torch.uint8), _check(params, torch.int32), _check(out)
    n, Hs, Ws, four = rgba.shape
    assert four == 4 and params.shape == (n, 6) and out.shape[0] == n and out.shape[1] == 3
    _work(0.0, (rgba, out))
    _call('r2n2_preprocess_views', _ptr(rgba), _ptr(params), _ptr(out), n, Hs, Ws, out.shape[2], out.shape[3],
          _DT[out.dtype], _stream())
    return out


def average_precision(pred, gt):
    """lib/test_net.py:69-70 on the GPU: float64 tensor [B] of per-sample average precision."""
    _check(pred, torch.float32), _check(gt, torch.float32)
    B, nv = pred.shape[0], pred.shape[1]
    ap = torch.empty(B, dtype=torch.float64, device=pred.device)
    _call('r2n2_average_precision', _ptr(pred), _ptr(gt), _ptr(ap), B, nv, _stream())
    return ap


def voxel_labels(occ, out=None):
    """occ uint8 [B,nv,nv,nv] (0 empty, 1 occupied, other = unfilled slot) -> y fp32 [B,nv,2,nv,nv] (lib/data_process.py:135-154)."""
    _check(occ, torch.uint8)
    B, nv = occ.shape[0], occ.shape[1]
    assert occ.shape == (B, nv, nv, nv)
    if out is None:
        out = torch.empty(B, nv, 2, nv, nv, dtype=torch.float32, device=occ.device)
    _check(out, torch.float32)
    assert out.shape == (B, nv, 2, nv, nv)
    _work(0.0, (occ, out))
    _call('r2n2_voxel_labels', _ptr(occ), _ptr(out), B, nv, _stream())
    return out


# ======================================================================= tape machinery
class Act:
    """An activation on the tape: channels-last data + gradient bookkeeping."""
    __slots__ = ('data', 'geom', 'grad', 'n_cons', 'n_recv', 'lrelu_out', 'grad_preact',
                 'grad_shared', 'requires_grad', 'useful', 'bias', 'bias_done')

    def __init__(self, data, geom, requires_grad=False, lrelu_out=False):
        self.data = data
        self.geom = tuple(int(g) for g in geom)   # (N, D, H, W, C)
        self.grad = None
        self.n_cons = 0          # consumers registered in the forward pass
        self.n_recv = 0          # gradient contributions received so far in the backward pass
        self.lrelu_out = lrelu_out      # data = LeakyReLU(z): sign(data) == sign(z)
        self.grad_preact = False        # grad already multiplied by LeakyReLU'(z)
        self.grad_shared = False        # grad buffer aliased by another Act: no in-place edits
        self.requires_grad = requires_grad
        self.useful = 1.0               # non-structural-zero fraction (1/8 after Unpool3D)
        self.bias = None                # bias Param of the conv that produced this activation
        self.bias_done = False          # its gradient was already formed by the epilogue that wrote .grad


class Param:
    """A packed parameter tensor used as a GEMM operand.
    data: fp32 master [cout, taps, cin] (contiguous view into the flat store)
    grad: fp32, same shape.  mirror: compute-dtype copy (bf16 path) or None.
    wt:   [cin, taps(reversed), cout] compute-dtype copy for the data-gradient GEMM."""

    def __init__(self, data, grad, mirror, cout, taps, cin, store=None, mirror_lo=None):
        self.data, self.grad, self.mirror = data, grad, mirror
        self.mirror_lo = mirror_lo      # bf16x3: low half of the operand split (mirror = high half)
        self.cout, self.taps, self.cin = cout, taps, cin
        self.store = store
        self.wt = None
        self.wt_lo = None
        self.tbatch_lo = None
        self.wt_version = -1
        self.tbatch = None       # TransposeBatch: all transposed operands refreshed by one launch
        self.grad_written = False

    def operand(self):
        if self.mirror_lo is not None:
            return SplitPair(self.mirror, self.mirror_lo, self.data)
        return self.mirror if self.mirror is not None else self.data

    def operand_t(self):
        if self.tbatch is not None:
            self.tbatch.ensure()
            if self.tbatch_lo is not None:
                self.tbatch_lo.ensure()
                return SplitPair(self.wt, self.wt_lo)
            return self.wt
        ver = self.store.version if self.store is not None else 0
        n = self.cin * self.taps * self.cout
        if self.mirror_lo is not None:                       # bf16x3 without the batched transposes (host tests)
            if self.wt is None or self.wt_version != ver:
                t32 = torch.empty(n, dtype=torch.float32, device=self.data.device)
                weight_transpose(self.data, t32, self.cout, self.taps, self.cin)
                self.wt = torch.empty(n, dtype=torch.bfloat16, device=self.data.device)
                self.wt_lo = torch.empty(n, dtype=torch.bfloat16, device=self.data.device)
                split_bf16(t32, self.wt, self.wt_lo)
                self.wt_version = ver
            return SplitPair(self.wt, self.wt_lo)
        if self.wt is None or self.wt_version != ver:
            dt = self.mirror.dtype if self.mirror is not None else torch.float32
            if self.wt is None:
                self.wt = torch.empty(n, dtype=dt, device=self.data.device)
            weight_transpose(self.data, self.wt, self.cout, self.taps, self.cin)
            self.wt_version = ver
        return self.wt


class Ctx:
    """Execution context: compute dtype, conv algorithms, training flag and the tape."""

    def __init__(self, dtype=torch.float32, algo=L.ALGO_SIMT, wgrad_algo=None, training=False):
        self.dtype = dtype
        self.algo = algo
        self.wgrad_algo = algo if wgrad_algo is None else wgrad_algo
        self.training = training
        self.tape = []

    # ---- helpers -------------------------------------------------------------------------
    def empty(self, numel, like, dtype=None):
        return torch.empty(int(numel), dtype=dtype or self.dtype, device=like.device)

    def consume(self, x):
        if self.training and x.requires_grad:
            x.n_cons += 1

    def accumulate(self, x, g, shared=False):
        """Add gradient tensor g into x.grad (taking ownership of g when it is the first)."""
        if not x.requires_grad:
            return
        if x.n_recv == 0:
            x.grad, x.grad_shared = g, shared
        elif x.grad_shared:
            new = torch.empty_like(x.grad)
            add(x.grad, g, new)
            x.grad, x.grad_shared = new, False
        else:
            add(x.grad, g, x.grad)
            split_invalidate(x.grad)
        x.n_recv += 1

    def ensure_preact(self, out):
        """out.grad currently holds dL/d(out); turn it into dL/dz where out = LeakyReLU(z)."""
        if out.lrelu_out and not out.grad_preact:
            dst = torch.empty_like(out.grad) if out.grad_shared else out.grad
            # dst is the complete pre-activation gradient: its column sums are the bias gradient of the convolution
            # that produced `out` -- taken in the same pass instead of a bias_grad pass over dst
            C = out.geom[4]
            b = getattr(out, 'bias', None)
            if (FUSE_BIAS_GRAD and FUSE_LRELU_BIAS and b is not None and not out.bias_done and b.grad_written
                    and C <= 512 and C % (8 if out.data.dtype == torch.bfloat16 else 4) == 0 and b.grad.numel() == C):
                lrelu_bwd(out.data, out.grad, None, dst, colsum=b.grad, C=C)
                out.bias_done = True
            else:
                lrelu_bwd(out.data, out.grad, None, dst)
            split_invalidate(dst)
            out.grad, out.grad_shared, out.grad_preact = dst, False, True

    def dgrad_into(self, x, dy, W, k, extra=None, down2=False):
        """x.grad (+)= conv_T(dy; W), fusing the pending accumulation (RESIDUAL) and, when this is
        the last contribution to a LeakyReLU output, the LeakyReLU' mask (MASK).
        down2: dy lives on the x2 grid of an UP2 convolution; only its even positions reach x."""
        if not x.requires_grad:
            return
        N, D, H, Wd, Cin = x.geom
        flags = L.CONV_DOWN2 if down2 else 0
        res = None
        out_buf = None
        if x.n_recv > 0:
            res = x.grad
            out_buf = torch.empty_like(x.grad) if x.grad_shared else x.grad
            flags |= L.EPI_RESIDUAL
            if extra is not None:
                add(extra, res, extra)
                split_invalidate(extra)
                res = extra
        elif extra is not None:
            res = extra
            flags |= L.EPI_RESIDUAL
        if out_buf is None:
            out_buf = self.empty(x.data.numel(), x.data)
        last = (x.n_recv == x.n_cons - 1)
        mask = None
        if x.lrelu_out and last:
            flags |= L.EPI_MASK
            mask = x.data
            x.grad_preact = True
        # last contribution: out_buf is the final gradient of x's pre-activation, so its column sums are the
        # bias gradient of the convolution that produced x (fused into the epilogue: no second pass)
        colsum = None
        if last and x.bias is not None and not x.bias_done and FUSE_BIAS_GRAD and x.bias.grad_written:
            colsum = x.bias.grad
            x.bias_done = True
        conv_fwd_x(dy, W.operand_t(), None, mask, res, out_buf, (N, D, H, Wd, W.cout), Cin, k,
                   flags, self.algo, colsum)
        split_invalidate(out_buf)
        x.grad, x.grad_shared = out_buf, False
        x.n_recv += 1

    def weight_grads(self, x_data, dy, geom, W, b, k, up2=False):
        conv_wgrad_x(x_data, dy, W.grad, geom, W.cout, k, self.wgrad_algo, W.grad_written, up2)
        W.grad_written = True
        if b is not None:
            bias_grad(dy, b.grad, W.cout, b.grad_written)
            b.grad_written = True

    def backward(self):
        for fn in reversed(self.tape):
            fn()
        self.tape = []

    # ---- differentiable ops --------------------------------------------------------------
    def zero_skip(self, cin, cout):
        """Unpool3D+Conv3D can run fused (UP2: only the taps that meet a stored value)."""
        return (self.algo in (L.ALGO_UMMA, L.ALGO_SPLIT) and umma_ok(cin, cout)
                and (self.dtype == torch.bfloat16 or self.algo == L.ALGO_SPLIT or ZERO_SKIP_ANY_DTYPE))

    def conv(self, x, W, b, k, lrelu=False, residual=None, out_dtype=None, up2=False, out_data=None, bias_via_ones=None):
        """y = [lrelu](conv(x; W) + b) [+ residual].  x: Act; W,b: Param; k=(kd,kh,kw).
        up2: the convolution runs over Unpool3D(x) (lib/layers.py:375-385) without materialising
        the 7/8-zero tensor; y lives on the x2 grid and the gradient reaches x directly.
        out_data: the output was already computed into this buffer (chunked head of the step, overlapped
        with the host-to-device copy): only the tape entry is made.
        bias_via_ones=(tap, channel): input channel `channel` is the constant 1 (r2n2_nchw_to_rowpack) and its weights
        are zero, so dW[:, tap, channel] of the CENTRE tap is sum_p dy[p, :] -- the bias gradient comes out of the
        weight-gradient GEMM and no column-sum pass over dy (557 MB for conv1a) is needed."""
        N, D, H, Wd, Cin = x.geom
        assert Cin == W.cin and k[0] * k[1] * k[2] == W.taps, (x.geom, W.cin, W.taps, k)
        assert not (lrelu and residual is not None)
        cout = W.cout
        if up2 and not self.zero_skip(Cin, cout):
            return self.conv(self.unpool(x), W, b, k, lrelu, residual, out_dtype)
        og = (N, 2 * D, 2 * H, 2 * Wd, cout) if up2 else (N, D, H, Wd, cout)
        global useful_frac
        if out_data is not None:
            y = out_data
        else:
            y = self.empty(og[0] * og[1] * og[2] * og[3] * cout, x.data, out_dtype)
            flags = (L.EPI_BIAS if b is not None else 0) | (L.EPI_LRELU if lrelu else 0)
            if residual is not None:
                flags |= L.EPI_RESIDUAL
            if up2:
                flags |= L.CONV_UP2
            useful_frac = x.useful
            conv_fwd_x(x.data, W.operand(), None if b is None else b.data, None,
                       None if residual is None else residual.data, y, x.geom, cout, k, flags, self.algo)
            useful_frac = 1.0
        out = Act(y, og, requires_grad=self.training, lrelu_out=lrelu)
        out.bias = b
        if bias_via_ones is not None and b is not None:
            out.bias_done = True          # neither a fused column sum nor a bias_grad pass: see bwd()
        if self.training:
            self.consume(x)
            if residual is not None:
                self.consume(residual)

            def bwd():
                if out.grad is None:
                    return
                self.ensure_preact(out)
                dy = out.grad
                global useful_frac
                useful_frac = x.useful
                self.weight_grads(x.data, dy, x.geom, W, None if out.bias_done else b, k, up2)
                if bias_via_ones is not None and b is not None:
                    tap, ch = bias_via_ones
                    gw = W.grad.view(W.cout, W.taps, W.cin)
                    if b.grad_written:
                        b.grad.add_(gw[:, tap, ch])
                    else:
                        b.grad.copy_(gw[:, tap, ch])
                        b.grad_written = True
                    gw[:, :, ch].zero_()          # the slot's weights are structural zeros and must stay zero
                self.dgrad_into(x, dy, W, k, down2=up2)
                useful_frac = 1.0
                if residual is not None:
                    self.accumulate(residual, dy)   # hand the buffer over (we are done with it)
                out.grad = None
            self.tape.append(bwd)
        return out

    def pool(self, a, b=None, out_data=None, out_code=None):
        """2x2/2 max-pool (pad 1) of a (+ b).  out_data (+ out_code): already computed (chunked head).
        Training: the forward pass records one code byte per output element (which window positions attain the
        maximum, sign of the rectified operand) and the backward pass runs from dy + code alone (POOL_CODES)."""
        N, D, H, W, C = a.geom
        assert D == 1
        Ho, Wo = H // 2 + 1, W // 2 + 1
        code = None
        if out_data is not None:
            y, code = out_data, out_code
        else:
            y = self.empty(N * Ho * Wo * C, a.data)
            if self.training and POOL_CODES:
                code = torch.empty(N * Ho * Wo * C, dtype=torch.uint8, device=a.data.device)
            maxpool_fwd(a.data, None if b is None else b.data, y, N, H, W, C, code)
        out = Act(y, (N, 1, Ho, Wo, C), requires_grad=self.training,
                  lrelu_out=(b is None and a.lrelu_out))
        if self.training:
            self.consume(a)
            if b is not None:
                self.consume(b)

            def bwd():
                if out.grad is None:
                    return
                # max commutes with the monotone LeakyReLU, so pool(lrelu(z)) passes the mask on
                fuse = (b is None and a.lrelu_out and a.n_recv == 0 and a.n_cons == 1
                        and not out.grad_preact)
                dx = torch.empty_like(a.data)

                def bias_sum(x, final):
                    # dx / dxb is the complete pre-activation gradient of x: its column sums are the bias
                    # gradient of the convolution that produced x (no separate pass)
                    if (final and FUSE_POOL_BWD and FUSE_BIAS_GRAD and x.bias is not None and not x.bias_done
                            and x.bias.grad_written and x.n_recv == 0 and x.n_cons == 1):
                        x.bias_done = True
                        return x.bias.grad
                    return None
                if b is None:
                    preact = fuse or out.grad_preact or not a.lrelu_out
                    if code is not None:
                        maxpool_bwd_code(code, out.grad, dx, N, H, W, C, fuse, colsum_a=bias_sum(a, preact))
                    else:
                        maxpool_bwd(a.data, None, out.grad, dx, N, H, W, C, fuse, colsum_a=bias_sum(a, preact))
                    if out.grad_preact:         # mask already applied downstream (on the max)
                        assert a.n_cons == 1
                        a.grad_preact = True
                    elif fuse:
                        a.grad_preact = True
                    self.accumulate(a, dx)
                else:
                    # pool(a + b): both branches receive dx; a rectified branch b that feeds nothing else gets its
                    # pre-activation gradient dx * LeakyReLU'(b) from the same kernel
                    split = (FUSE_POOL_BWD and b.requires_grad and b.lrelu_out and not b.grad_preact
                             and b.n_recv == 0 and b.n_cons == 1)
                    dxb = torch.empty_like(b.data) if split else None
                    sums = dict(colsum_a=bias_sum(a, not a.lrelu_out), colsum_b=bias_sum(b, True) if split else None)
                    if code is not None:
                        maxpool_bwd_code(code, out.grad, dx, N, H, W, C, False, dxb=dxb, **sums)
                    else:
                        maxpool_bwd(a.data, b.data, out.grad, dx, N, H, W, C, False, dxb=dxb, **sums)
                    if split:
                        self.accumulate(a, dx)
                        b.grad_preact = True
                        self.accumulate(b, dxb)
                    else:
                        self.accumulate(a, dx, shared=b.requires_grad)
                        self.accumulate(b, dx, shared=a.requires_grad)
                out.grad = None
            self.tape.append(bwd)
        return out

    def unpool(self, a, b=None):
        """Unpool3D x2 (zero stuffing) of a (+ b)."""
        B, D, H, W, C = a.geom
        y = self.empty(B * D * H * W * 8 * C, a.data)
        unpool_fwd(a.data, None if b is None else b.data, y, B, D, H, W, C)
        out = Act(y, (B, 2 * D, 2 * H, 2 * W, C), requires_grad=self.training)
        out.useful = 0.125          # 7/8 structural zeros: only 1/8 of a conv's MACs are useful
        if self.training:
            self.consume(a)
            if b is not None:
                self.consume(b)

            def bwd():
                if out.grad is None:
                    return
                dx = torch.empty_like(a.data)
                unpool_bwd(out.grad, dx, B, D, H, W, C)
                if b is None:
                    self.accumulate(a, dx)
                else:
                    self.accumulate(a, dx, shared=b.requires_grad)
                    self.accumulate(b, dx, shared=a.requires_grad)
                out.grad = None
            self.tape.append(bwd)
        return out

    def reshape(self, x, geom):
        """FlattenLayer: reinterpret the (contiguous, channels-last) buffer with a new geometry."""
        geom = tuple(int(g) for g in geom)
        assert x.n_cons == 0 and x.data.numel() == geom[0] * geom[1] * geom[2] * geom[3] * geom[4]
        x.geom = geom
        return x

    def add_act(self, a, b):
        """AddLayer as a standalone op (only used by the layer-by-layer path)."""
        y = torch.empty_like(a.data)
        add(a.data, b.data, y)
        out = Act(y, a.geom, requires_grad=self.training)
        if self.training:
            self.consume(a), self.consume(b)

            def bwd():
                if out.grad is None:
                    return
                self.accumulate(a, out.grad, shared=b.requires_grad)
                self.accumulate(b, out.grad, shared=a.requires_grad)
                out.grad = None
            self.tape.append(bwd)
        return out

    def lrelu(self, x):
        """Standalone LeakyReLU (layer-by-layer path; the fused path folds it into conv)."""
        y = torch.empty_like(x.data)
        lrelu_fwd(x.data, y)
        out = Act(y, x.geom, requires_grad=self.training, lrelu_out=True)
        if self.training:
            self.consume(x)

            def bwd():
                if out.grad is None:
                    return
                self.ensure_preact(out)
                self.accumulate(x, out.grad)
                out.grad = None
            self.tape.append(bwd)
        return out

    # ---- ops of the layer-by-layer API (user-defined network_definition; models/net.py:56-58) -----------
    def view(self, x, geom):
        """The same buffer under another geometry (FlattenLayer, the (B,1,1,1,n) input of a dot, the
        reshape of dot(f, Wx) onto the hidden grid): a new Act whose gradient flows back to x."""
        geom = tuple(int(g) for g in geom)
        assert x.data.numel() == geom[0] * geom[1] * geom[2] * geom[3] * geom[4], (x.geom, geom)
        out = Act(x.data, geom, requires_grad=self.training and x.requires_grad)
        out.useful = x.useful
        if out.requires_grad:
            self.consume(x)

            def bwd():
                if out.grad is None:
                    return
                self.accumulate(x, out.grad)
                out.grad = None
            self.tape.append(bwd)
        return out

    def _unary(self, x, op, bwd_op):
        y = torch.empty_like(x.data)
        eltwise(op, x.data, None, y)
        out = Act(y, x.geom, requires_grad=self.training and x.requires_grad)
        if out.requires_grad:
            self.consume(x)

            def bwd():
                if out.grad is None:
                    return
                dx = torch.empty_like(y)
                if bwd_op == L.ELT_NEG:
                    eltwise(L.ELT_NEG, out.grad, None, dx)
                else:
                    eltwise(bwd_op, y, out.grad, dx)
                self.accumulate(x, dx)
                out.grad = None
            self.tape.append(bwd)
        return out

    def sigmoid(self, x):
        """SigmoidLayer (lib/layers.py:649-656)"""
        return self._unary(x, L.ELT_SIGMOID, L.ELT_SIGMOID_BWD)

    def tanh(self, x):
        """TanhLayer (lib/layers.py:659-665)"""
        return self._unary(x, L.ELT_TANH, L.ELT_TANH_BWD)

    def complement(self, x):
        """ComplementLayer (lib/layers.py:668-676): 1 - x"""
        return self._unary(x, L.ELT_COMPLEMENT, L.ELT_NEG)

    def mul(self, a, b):
        """EltwiseMultiplyLayer (lib/layers.py:217-225)"""
        assert a.data.numel() == b.data.numel()
        y = torch.empty_like(a.data)
        eltwise(L.ELT_MUL, a.data, b.data, y)
        out = Act(y, a.geom, requires_grad=self.training and (a.requires_grad or b.requires_grad))
        if out.requires_grad:
            self.consume(a), self.consume(b)

            def bwd():
                if out.grad is None:
                    return
                for x, other in ((a, b), (b, a)):
                    if x.requires_grad:
                        dx = torch.empty_like(y)
                        eltwise(L.ELT_MUL, out.grad, other.data, dx)
                        self.accumulate(x, dx)
                out.grad = None
            self.tape.append(bwd)
        return out

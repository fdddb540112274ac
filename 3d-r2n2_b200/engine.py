"""Fused executor for the three networks of the hot path (GRUNet, ResidualGRUNet and the
3D-conv-LSTM variant of the residual net).

Same arithmetic as the reference graphs (models/gru_net.py:63-144, models/res_gru_net.py:78-204),
re-scheduled B200-first:
  * everything that does not depend on h_{t-1} (the 2-D encoder, fc7 and the FCConv3DLayer
    x-projections dot(f, Wx), lib/layers.py:502-503) is hoisted out of the theano.scan recurrence
    and run ONCE over all T*B views;
  * the update/reset gates share one 3-D convolution of h (Cout=256) and one x-projection GEMM;
  * AddLayer + PoolLayer / AddLayer + Unpool3DLayer pairs are single kernels; bias + LeakyReLU
    (+ residual add, + LeakyReLU' in the backward) live in the GEMM epilogues;
  * all parameters, gradients and Adam moments live in ONE flat fp32 buffer each (packed
    layouts), so the optimiser is one launch and the data-parallel all-reduce a few big buckets.
"""
import math

import numpy as np
import os

import torch

from . import _lib as L
from . import ops
from . import packing as pk
from .ops import Act, Param, Ctx

N_CONVFILTER = [96, 128, 256, 256, 256, 256]
N_FC = 1024
N_DECONV = [128, 128, 128, 64, 32, 2]
N_GRU_VOX = 4
HID_POS = N_GRU_VOX ** 3          # 64 cells
HID_C = N_DECONV[0]               # 128 channels
# conv-GRU gate math in the epilogue of the hidden convolutions (tcgen05 path); R2N2_NO_GRU_FUSE=1: separate gate kernels
FUSE_GRU_EPILOGUE = not os.environ.get('R2N2_NO_GRU_FUSE')

NETS = ('GRUNet', 'ResidualGRUNet', 'ResidualLSTMNet')


def _round_up(n, m):
    return (n + m - 1) // m * m


class RefWeight:
    """One parameter tensor of the reference's `net.params` list (weights.npy order):
    knows how to move between the reference layout and its packed view in the flat store."""

    def __init__(self, name, ref_shape, is_bias, view, to_packed, to_ref):
        self.name, self.ref_shape, self.is_bias = name, tuple(ref_shape), is_bias
        self.view = view              # fp32 torch view (possibly strided) into the flat store
        self._to_packed, self._to_ref = to_packed, to_ref

    def set_value(self, arr):
        t = torch.as_tensor(np.asarray(arr, dtype=np.float32))
        if tuple(t.shape) != self.ref_shape:
            raise ValueError('%s: expected shape %s, got %s' % (self.name, self.ref_shape, tuple(t.shape)))
        self.view.copy_(self._to_packed(t.to(self.view.device)).reshape(self.view.shape))

    def get_value(self, src=None):
        v = self.view if src is None else src
        return self._to_ref(v).reshape(self.ref_shape).cpu().numpy()


class ParamStore:
    """Flat fp32 parameter / gradient / moment buffers.  Non-bias segments first (L2 decay
    applies to [0, n_decay)), biases after.  Segments are 64-element aligned."""

    def __init__(self, device, mirror_dtype=None, split=False):
        self.device = device
        self.mirror_dtype = mirror_dtype
        self.split = split            # bf16x3: mirror = hi, mirror_lo = lo operand halves
        self.mirror_lo = None
        self._plan = []           # (name, numel, is_bias)
        self.offsets = {}
        self.version = 0
        self.p = self.g = self.m = self.v = self.mirror = None

    def plan(self, name, numel, is_bias):
        self._plan.append((name, int(numel), bool(is_bias)))

    def allocate(self):
        off = 0
        for want_bias in (False, True):
            for name, numel, is_bias in self._plan:
                if is_bias == want_bias:
                    self.offsets[name] = (off, numel)
                    off += _round_up(numel, 64)
            if not want_bias:
                self.n_decay = off
        self.n = off
        self.p = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        self.g = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        if self.mirror_dtype is not None:
            self.mirror = torch.zeros(self.n, dtype=self.mirror_dtype, device=self.device)
            if self.split:
                self.mirror_lo = torch.zeros(self.n, dtype=self.mirror_dtype, device=self.device)

    def seg(self, name, buf=None):
        off, numel = self.offsets[name]
        return (self.p if buf is None else buf)[off:off + numel]

    def ensure_moments(self):
        if self.m is None:
            self.m = torch.zeros_like(self.p)
            self.v = torch.zeros_like(self.p)

    def touch(self):
        """Parameters changed (set_value / optimiser step): refresh the operand mirrors lazily."""
        self.version += 1

    def refresh_mirror(self):
        if self.mirror_lo is not None:
            ops.split_bf16(self.p, self.mirror, self.mirror_lo)
        elif self.mirror is not None and self.p.is_cuda:
            ops.cast(self.p, self.mirror)


class Engine:
    def __init__(self, net_name, batch_size, device='cuda', compute='fp32', conv_algo=None,
                 wgrad_algo=None, img=127, n_vox=32, params_only=False):
        if net_name not in NETS:
            raise ValueError('unknown network %r' % (net_name,))
        self.params_only = params_only      # host-side parameter bookkeeping only (CPU tests)
        if not params_only:
            if not torch.cuda.is_available():
                raise RuntimeError('r2n2_b200.Engine needs a CUDA device (no CPU fallback exists)')
            L.load()
        self.net_name = net_name
        self.B = int(batch_size)
        self.img, self.n_vox = img, n_vox
        self.device = torch.device(device)
        self.compute = compute
        self._graphs = {}
        # 'fp32': FFMA kernels (exact parity path); 'bf16': tcgen05, bf16 operands and activations;
        # 'bf16x3': tcgen05 over hi/lo split operands, fp32 activations (fp32 accuracy at 1/3 of the bf16 rate)
        self.dtype = {'fp32': torch.float32, 'bf16': torch.bfloat16, 'bf16x3': torch.float32}[compute]
        self.split = compute == 'bf16x3'
        if conv_algo is None:
            conv_algo = {'fp32': L.ALGO_SIMT, 'bf16': L.ALGO_UMMA, 'bf16x3': L.ALGO_SPLIT}[compute]
        self.algo = conv_algo
        self.wgrad_algo = conv_algo if wgrad_algo is None else wgrad_algo
        self.residual = net_name != 'GRUNet'
        self.lstm = net_name == 'ResidualLSTMNet'
        self.n_gates = 3
        # channel padding of the RGB input: 4 for the fp32 SIMT path, 16 for tensor-core K tiles
        self.cin_pad = 4 if self.algo == L.ALGO_SIMT else 16
        # conv11 has 2 output channels; on the tensor-core path it is padded to one N granule (16):
        # the extra weight rows are zero and stay zero (their gradients are exactly zero)
        self.logit_c = 2 if self.algo == L.ALGO_SIMT else 16
        # tensor path: the 7x7x3 first layer runs on a row-packed input (7 taps x 3 channels packed
        # into 32 channels by r2n2_nchw_to_rowpack) as a 7x1 convolution: K = 7*32 instead of 49*16
        self.rowpack = 0 if self.algo == L.ALGO_SIMT else 32
        self.store = ParamStore(self.device, torch.bfloat16 if self.split else
                                (None if self.dtype == torch.float32 else self.dtype), split=self.split)
        self.P = {}
        self.weights = []      # RefWeight list in reference creation order
        self._define()
        self.ctx = None
        self.last = None
        self.progress = None       # callback(stage) fired when a backward stage's grads are final
        self._marks = (0, 0)
        self._enc_marks = {}       # encoder sub-stage -> tape index at which its gradients are final
        self.debug = None          # dict: forward() drops named intermediate buffers here (tools/diag_head.py)

    # ------------------------------------------------------------------ parameter definition
    def _define(self):
        c, d = N_CONVFILTER, N_DECONV
        st = self.store
        defs = []        # deferred (after allocate) construction of Params / RefWeights

        def conv2(name, cout, cin, k, cin_pad=None):
            cp = cin_pad or cin
            if cin_pad and self.rowpack:            # first layer, row-packed: [cout][k][rowpack]
                st.plan(name + '.W', cout * k * self.rowpack, False)
                st.plan(name + '.b', cout, True)
                defs.append(('conv2rp', name, cout, cin, k, self.rowpack))
                return
            st.plan(name + '.W', cout * k * k * cp, False)
            st.plan(name + '.b', cout, True)
            defs.append(('conv2', name, cout, cin, k, cp))

        def conv3(name, cout, cin, k):
            cs = self.logit_c if name == 'conv11' else cout      # stored (padded) out channels
            st.plan(name + '.W', cs * k ** 3 * cin, False)
            st.plan(name + '.b', cs, True)
            defs.append(('conv3', name, cout, cin, k, cs))

        if self.residual:
            conv2('conv1a', c[0], 3, 7, self.cin_pad); conv2('conv1b', c[0], c[0], 3)
            conv2('conv2a', c[1], c[0], 3); conv2('conv2b', c[1], c[1], 3); conv2('conv2c', c[1], c[0], 1)
            conv2('conv3a', c[2], c[1], 3); conv2('conv3b', c[2], c[2], 3); conv2('conv3c', c[2], c[1], 1)
            conv2('conv4a', c[3], c[2], 3); conv2('conv4b', c[3], c[3], 3)
            conv2('conv5a', c[4], c[3], 3); conv2('conv5b', c[4], c[4], 3); conv2('conv5c', c[4], c[3], 1)
            conv2('conv6a', c[5], c[4], 3); conv2('conv6b', c[5], c[5], 3)
        else:
            cin = 3
            for i in range(6):
                conv2('conv%d' % (i + 1), c[i], cin, 7 if i == 0 else 3, self.cin_pad if i == 0 else None)
                cin = c[i]
        st.plan('fc7.W', N_FC * c[5] * 9, False)
        st.plan('fc7.b', N_FC, True)
        defs.append(('fc7',))
        # recurrent unit: gate groups that read the same input share one conv / one GEMM
        if self.lstm:
            self.gate_names = ['t_x_s_forget', 't_x_s_input', 't_x_s_cell']
            self.groups = [('fig', [0, 1, 2])]
        else:
            self.gate_names = ['t_x_s_update', 't_x_s_reset', 't_x_rs']
            self.groups = [('ur', [0, 1]), ('c', [2])]
        for gname, members in self.groups:
            ng = len(members)
            st.plan('rnn.Wh_' + gname, ng * HID_C * 27 * HID_C, False)
            st.plan('rnn.Wx_' + gname, HID_POS * ng * HID_C * N_FC, False)
            st.plan('rnn.b_' + gname, ng * HID_C, True)
        defs.append(('rnn',))
        if self.residual:
            conv3('conv7a', d[1], d[0], 3); conv3('conv7b', d[1], d[1], 3)
            conv3('conv8a', d[2], d[1], 3); conv3('conv8b', d[2], d[2], 3)
            conv3('conv9a', d[3], d[2], 3); conv3('conv9b', d[3], d[3], 3); conv3('conv9c', d[3], d[2], 1)
            conv3('conv10a', d[4], d[3], 3); conv3('conv10b', d[4], d[4], 3); conv3('conv10c', d[4], d[4], 3)
            conv3('conv11', d[5], d[4], 3)
        else:
            conv3('conv7', d[1], d[0], 3); conv3('conv8', d[2], d[1], 3); conv3('conv9', d[3], d[2], 3)
            conv3('conv10', d[4], d[3], 3); conv3('conv11', d[5], d[4], 3)
        st.allocate()

        def mk_param(seg, cout, taps, cin):
            p = Param(st.seg(seg), st.seg(seg, st.g),
                      None if st.mirror is None else st.seg(seg, st.mirror), cout, taps, cin, st,
                      None if st.mirror_lo is None else st.seg(seg, st.mirror_lo))
            self.P[seg] = p
            return p

        def mk_bias(seg, n):
            p = Param(st.seg(seg), st.seg(seg, st.g), None, n, 1, 1, st)
            self.P[seg] = p
            return p

        for dfn in defs:
            kind = dfn[0]
            if kind == 'conv2':
                _, name, cout, cin, k, cp = dfn
                mk_param(name + '.W', cout, k * k, cp)
                mk_bias(name + '.b', cout)
                vw = st.seg(name + '.W').view(cout, k, k, cp)
                self.weights.append(RefWeight(
                    name + '.W', (cout, cin, k, k), False, vw,
                    lambda t, cp=cp: pk.pack_conv2d(t, cp), lambda v, cin=cin: pk.unpack_conv2d(v, cin)))
                self.weights.append(RefWeight(name + '.b', (cout,), True, st.seg(name + '.b'),
                                              lambda t: t, lambda v: v))
            elif kind == 'conv2rp':
                _, name, cout, cin, k, rp = dfn
                mk_param(name + '.W', cout, k, rp)
                mk_bias(name + '.b', cout)
                vw = st.seg(name + '.W').view(cout, k, rp)
                self.weights.append(RefWeight(
                    name + '.W', (cout, cin, k, k), False, vw,
                    lambda t, rp=rp: pk.pack_conv2d_rowpack(t, rp),
                    lambda v, cin=cin, k=k: pk.unpack_conv2d_rowpack(v, cin, k)))
                self.weights.append(RefWeight(name + '.b', (cout,), True, st.seg(name + '.b'),
                                              lambda t: t, lambda v: v))
            elif kind == 'conv3':
                _, name, cout, cin, k, cs = dfn
                mk_param(name + '.W', cs, k ** 3, cin)
                mk_bias(name + '.b', cs)
                vw = st.seg(name + '.W').view(cs, k, k, k, cin)[:cout]
                self.weights.append(RefWeight(name + '.W', (cout, k, cin, k, k), False, vw,
                                              pk.pack_conv3d, pk.unpack_conv3d))
                self.weights.append(RefWeight(name + '.b', (cout,), True, st.seg(name + '.b')[:cout],
                                              lambda t: t, lambda v: v))
            elif kind == 'fc7':
                n_in = N_CONVFILTER[5] * 9
                mk_param('fc7.W', N_FC, 1, n_in)
                mk_bias('fc7.b', N_FC)
                chw = (N_CONVFILTER[5], 3, 3)
                self.weights.append(RefWeight('fc7.W', (n_in, N_FC), False,
                                              st.seg('fc7.W').view(N_FC, n_in),
                                              lambda t: pk.pack_fc(t, chw), lambda v: pk.unpack_fc(v, chw)))
                self.weights.append(RefWeight('fc7.b', (N_FC,), True, st.seg('fc7.b'),
                                              lambda t: t, lambda v: v))
            elif kind == 'rnn':
                slot = {}
                for gname, members in self.groups:
                    ng = len(members)
                    mk_param('rnn.Wh_' + gname, ng * HID_C, 27, HID_C)
                    mk_param('rnn.Wx_' + gname, HID_POS * ng * HID_C, 1, N_FC)
                    mk_bias('rnn.b_' + gname, ng * HID_C)
                    wh = st.seg('rnn.Wh_' + gname).view(ng, HID_C, 3, 3, 3, HID_C)
                    wx = st.seg('rnn.Wx_' + gname).view(HID_POS, ng, HID_C, N_FC)
                    bb = st.seg('rnn.b_' + gname).view(ng, HID_C)
                    for j, gi in enumerate(members):
                        slot[gi] = (wh[j], wx[:, j], bb[j])
                g4 = N_GRU_VOX
                for gi, lname in enumerate(self.gate_names):
                    wh, wx, bb = slot[gi]
                    self.weights.append(RefWeight(lname + '.Wh', (HID_C, 3, HID_C, 3, 3), False, wh,
                                                  pk.pack_conv3d, pk.unpack_conv3d))
                    self.weights.append(RefWeight(
                        lname + '.Wx', (N_FC, HID_POS * HID_C), False, wx,
                        lambda t: pk.pack_fcconv_wx(t, g4, HID_C, g4, g4),
                        lambda v: pk.unpack_fcconv_wx(v, g4, HID_C, g4, g4)))
                    self.weights.append(RefWeight(lname + '.b', (HID_C,), True, bb,
                                                  lambda t: t, lambda v: v))

        if st.p.is_cuda:
            # conv1's input has no gradient: its transposed operand is never used
            tparams = [p for n, p in self.P.items() if p.taps * p.cin > 1 and not n.startswith(('conv1a.', 'conv1.'))]
            self.tbatch = ops.TransposeBatch(tparams, st)
            if self.split:
                self.tbatch_lo = ops.TransposeBatch(tparams, st, 'mirror_lo', 'wt_lo')

    # ------------------------------------------------------------------ parameter I/O
    def set_params(self, arrays):
        """arrays: list in the reference's weights.npy order (models/net.py:60-86)."""
        if len(arrays) != len(self.weights):
            raise ValueError('expected %d parameter tensors, got %d' % (len(self.weights), len(arrays)))
        for w, a in zip(self.weights, arrays):
            w.set_value(a)
        self.params_changed()

    def get_params(self):
        return [w.get_value() for w in self.weights]

    def get_grads(self):
        """Gradients in reference layout / order (tensor.grad(loss, params), models/net.py:58)."""
        out = []
        for w in self.weights:
            off = w.view.storage_offset() - self.store.p.storage_offset()
            gview = torch.as_strided(self.store.g, w.view.shape, w.view.stride(), off)
            out.append(w.get_value(gview))
        return out

    def params_changed(self):
        self.store.touch()
        self.store.refresh_mirror()

    # ------------------------------------------------------------------ recurrent composites
    def _xproj(self, ctx, f, TB):
        """Hoisted FCConv3DLayer x-projections (lib/layers.py:502-503) for all T*B views."""
        xps = {}
        for gname, members in self.groups:
            n = HID_POS * len(members) * HID_C
            xp = ctx.empty(TB * n, f.data)
            ops.conv_fwd_x(f.data, self.P['rnn.Wx_' + gname].operand(), None, None, None, xp,
                         (TB, 1, 1, 1, N_FC), n, (1, 1, 1), 0, ctx.algo)
            xps[gname] = xp
        return xps

    def _rnn_param_grads(self, ctx, f, dxp, hin, T, B):
        """Weight gradients of the recurrent unit, batched over all time steps, and dL/df."""
        TB = T * B
        g4 = N_GRU_VOX
        first = None
        for idx, (gname, members) in enumerate(self.groups):
            ng = len(members)
            cg = ng * HID_C
            Wh, Wx, bb = self.P['rnn.Wh_' + gname], self.P['rnn.Wx_' + gname], self.P['rnn.b_' + gname]
            d = dxp[gname]                                     # [T,B,64,cg]
            ops.bias_grad(d, bb.grad, cg, bb.grad_written); bb.grad_written = True
            if T > 1:                                          # step 0 has h == 0: no conv term
                ops.conv_wgrad_x(hin[gname][:(T - 1) * B * HID_POS * HID_C], d[B * HID_POS * cg:],
                               Wh.grad, ((T - 1) * B, g4, g4, g4, HID_C), cg, (3, 3, 3),
                               ctx.wgrad_algo, Wh.grad_written)
                Wh.grad_written = True
            ops.conv_wgrad_x(f.data, d, Wx.grad, (TB, 1, 1, 1, N_FC), HID_POS * cg, (1, 1, 1),
                           ctx.wgrad_algo, Wx.grad_written)
            Wx.grad_written = True
            if idx < len(self.groups) - 1:
                first = ctx.empty(TB * N_FC, f.data)
                ops.conv_fwd_x(d, Wx.operand_t(), None, None, None, first,
                             (TB, 1, 1, 1, HID_POS * cg), N_FC, (1, 1, 1), 0, ctx.algo)
            else:
                # last group: fuse the accumulation and fc7's LeakyReLU' into the epilogue
                self._dgrad_fc(ctx, f, d, Wx, HID_POS * cg, first)

    def _dgrad_fc(self, ctx, f, d, Wx, kdim, extra):
        TB = f.geom[0]
        flags, res, mask = 0, None, None
        if f.n_recv > 0:
            raise RuntimeError('fc7 output has an unexpected second consumer')
        if extra is not None:
            flags |= L.EPI_RESIDUAL
            res = extra
        if f.lrelu_out and f.n_cons == 1:
            flags |= L.EPI_MASK
            mask = f.data
            f.grad_preact = True
        out = extra if extra is not None else ctx.empty(TB * N_FC, f.data)
        ops.conv_fwd_x(d, Wx.operand_t(), None, mask, res, out, (TB, 1, 1, 1, kdim), N_FC, (1, 1, 1),
                     flags, ctx.algo)
        f.grad, f.grad_shared = out, False
        f.n_recv += 1

    def gru_sequence(self, ctx, f, T, B):
        """theano.scan over views of the GRU update (models/res_gru_net.py:128-160):
        u = sigmoid(conv(h;Wh_u) + f.Wx_u + b_u), r likewise, c = tanh(conv(r*h;Wh_c) + f.Wx_c + b_c),
        h <- u*h + (1-u)*c.   Returns (h_T Act, update gates [T,B,64,128])."""
        g4 = N_GRU_VOX
        hn = HID_POS * HID_C
        P = self.P
        xps = self._xproj(ctx, f, T * B)
        xp_ur, xp_c = xps['ur'], xps['c']
        u = ctx.empty(T * B * hn, f.data); r = ctx.empty(T * B * hn, f.data)
        rh = ctx.empty(T * B * hn, f.data); c = ctx.empty(T * B * hn, f.data)
        h = ctx.empty(T * B * hn, f.data)
        sl = lambda buf, t, width=hn: buf[t * B * width:(t + 1) * B * width]
        geom_h = (B, g4, g4, g4, HID_C)
        # tcgen05 path: the gate math runs in the epilogue of the hidden convolutions (r2n2_conv_fwd_gru) -- no conv
        # output tensor, no separate gate launches except for the first view (h = 0: nothing to convolve)
        # (ops.ZERO_SKIP_ANY_DTYPE: host-logic tests on the emulated ABI run this schedule in fp32 against the oracle)
        fused = (FUSE_GRU_EPILOGUE and ctx.algo == L.ALGO_UMMA
                 and (f.data.dtype == torch.bfloat16 or ops.ZERO_SKIP_ANY_DTYPE))
        for t in range(T):
            if t == 0:
                ops.gru_ur_fwd(None, sl(xp_ur, 0, 2 * hn), P['rnn.b_ur'].data, None, sl(u, 0), sl(r, 0),
                               sl(rh, 0), B)
                ops.gru_out_fwd(None, sl(xp_c, 0), P['rnn.b_c'].data, sl(u, 0), None, sl(c, 0),
                                sl(h, 0), B)
                continue
            hp = sl(h, t - 1)
            if fused:
                ops.conv_fwd_gru(L.GRU_UR, hp, P['rnn.Wh_ur'].operand(), P['rnn.b_ur'].data,
                                 (sl(xp_ur, t, 2 * hn), hp), (sl(u, t), sl(r, t), sl(rh, t)), geom_h, 2 * HID_C, (3, 3, 3))
                ops.conv_fwd_gru(L.GRU_BLEND, sl(rh, t), P['rnn.Wh_c'].operand(), P['rnn.b_c'].data,
                                 (sl(xp_c, t), sl(u, t), hp), (sl(c, t), sl(h, t)), geom_h, HID_C, (3, 3, 3))
                continue
            a_ur = ctx.empty(B * 2 * hn, f.data)
            ops.conv_fwd_x(hp, P['rnn.Wh_ur'].operand(), None, None, None, a_ur, geom_h, 2 * HID_C,
                         (3, 3, 3), 0, ctx.algo)
            ops.gru_ur_fwd(a_ur, sl(xp_ur, t, 2 * hn), P['rnn.b_ur'].data, hp, sl(u, t), sl(r, t),
                           sl(rh, t), B)
            a_c = ctx.empty(B * hn, f.data)
            ops.conv_fwd_x(sl(rh, t), P['rnn.Wh_c'].operand(), None, None, None, a_c, geom_h, HID_C,
                         (3, 3, 3), 0, ctx.algo)
            ops.gru_out_fwd(a_c, sl(xp_c, t), P['rnn.b_c'].data, sl(u, t), hp, sl(c, t), sl(h, t), B)
        h_last = Act(sl(h, T - 1), geom_h, requires_grad=ctx.training)
        if ctx.training:
            ctx.consume(f)

            def bwd():
                if h_last.grad is None:
                    return
                dxp_ur = torch.zeros(T * B * 2 * hn, dtype=ctx.dtype, device=f.data.device)
                dxp_c = ctx.empty(T * B * hn, f.data)
                dh = h_last.grad
                geom_ur = (B, g4, g4, g4, 2 * HID_C)
                out_bwd_done = False            # fused: step t's gate gradients came out of step t+1's last epilogue
                for t in range(T - 1, -1, -1):
                    hp = sl(h, t - 1) if t > 0 else None
                    if not out_bwd_done:
                        dh_prev = ctx.empty(B * hn, f.data)
                        ops.gru_out_bwd(dh, sl(u, t), sl(c, t), hp, sl(dxp_ur, t, 2 * hn), sl(dxp_c, t),
                                        dh_prev, B)
                    if t == 0:
                        break
                    if fused:
                        # d(r*h) = conv(d pre-tanh; Wh_c^T) with the reset-gate gradient in its epilogue, then
                        # dh_{t-1} = conv(d pre-sigmoid; Wh_ur^T) + dh_prev with step t-1's gate gradients in ITS epilogue
                        ops.conv_fwd_gru(L.GRU_RBWD, sl(dxp_c, t), P['rnn.Wh_c'].operand_t(), None, (hp, sl(r, t)),
                                         (sl(dxp_ur, t, 2 * hn), dh_prev), geom_h, HID_C, (3, 3, 3))
                        dh_next = ctx.empty(B * hn, f.data)
                        ops.conv_fwd_gru(L.GRU_OBWD, sl(dxp_ur, t, 2 * hn), P['rnn.Wh_ur'].operand_t(), None,
                                         (dh_prev, sl(u, t - 1), sl(c, t - 1), sl(h, t - 2) if t > 1 else None),
                                         (sl(dxp_c, t - 1), sl(dxp_ur, t - 1, 2 * hn), dh_next), geom_ur, HID_C, (3, 3, 3))
                        dh_prev = dh_next
                        out_bwd_done = True
                        continue
                    d_rh = ctx.empty(B * hn, f.data)
                    ops.conv_fwd_x(sl(dxp_c, t), P['rnn.Wh_c'].operand_t(), None, None, None, d_rh,
                                 geom_h, HID_C, (3, 3, 3), 0, ctx.algo)
                    ops.gru_r_bwd(d_rh, hp, sl(r, t), sl(dxp_ur, t, 2 * hn), dh_prev, B)
                    ops.conv_fwd_x(sl(dxp_ur, t, 2 * hn), P['rnn.Wh_ur'].operand_t(), None, None, dh_prev,
                                 dh_prev, geom_ur, HID_C, (3, 3, 3), L.EPI_RESIDUAL, ctx.algo)
                    dh = dh_prev
                self._rnn_param_grads(ctx, f, {'ur': dxp_ur, 'c': dxp_c}, {'ur': h, 'c': rh[B * hn:]},
                                      T, B)
                h_last.grad = None
            ctx.tape.append(bwd)
        return h_last, u

    def lstm_sequence(self, ctx, f, T, B):
        """3D-conv-LSTM variant (no reference code exists):
        f,i = sigmoid(.), g = tanh(.), all three FCConv3D terms read h_{t-1};
        s <- f*s + i*g ; h <- tanh(s).  Returns (h_T Act, input gates [T,B,64,128] strided)."""
        g4 = N_GRU_VOX
        hn = HID_POS * HID_C
        P = self.P
        xp = self._xproj(ctx, f, T * B)['fig']
        fig = ctx.empty(T * B * 3 * hn, f.data)
        s = ctx.empty(T * B * hn, f.data); h = ctx.empty(T * B * hn, f.data)
        sl = lambda buf, t, width=hn: buf[t * B * width:(t + 1) * B * width]
        geom_h = (B, g4, g4, g4, HID_C)
        for t in range(T):
            a = None
            if t > 0:
                a = ctx.empty(B * 3 * hn, f.data)
                ops.conv_fwd_x(sl(h, t - 1), P['rnn.Wh_fig'].operand(), None, None, None, a, geom_h,
                             3 * HID_C, (3, 3, 3), 0, ctx.algo)
            ops.lstm_fwd(a, sl(xp, t, 3 * hn), P['rnn.b_fig'].data, sl(s, t - 1) if t > 0 else None,
                         sl(fig, t, 3 * hn), sl(s, t), sl(h, t), B)
        h_last = Act(sl(h, T - 1), geom_h, requires_grad=ctx.training)
        if ctx.training:
            ctx.consume(f)

            def bwd():
                if h_last.grad is None:
                    return
                dxp = ctx.empty(T * B * 3 * hn, f.data)
                dh, ds = h_last.grad, None
                geom_3 = (B, g4, g4, g4, 3 * HID_C)
                for t in range(T - 1, -1, -1):
                    ds_prev = ctx.empty(B * hn, f.data)
                    ops.lstm_bwd(dh, ds, sl(fig, t, 3 * hn), sl(s, t - 1) if t > 0 else None, sl(h, t),
                                 sl(dxp, t, 3 * hn), ds_prev, B)
                    if t == 0:
                        break
                    dh = ctx.empty(B * hn, f.data)
                    ops.conv_fwd_x(sl(dxp, t, 3 * hn), P['rnn.Wh_fig'].operand_t(), None, None, None, dh,
                                 geom_3, HID_C, (3, 3, 3), 0, ctx.algo)
                    ds = ds_prev
                self._rnn_param_grads(ctx, f, {'fig': dxp}, {'fig': h}, T, B)
                h_last.grad = None
            ctx.tape.append(bwd)
        gates_i = fig.view(T, B, HID_POS, 3, HID_C)[:, :, :, 1]
        return h_last, gates_i

    # ------------------------------------------------------------------ whole network
    def forward(self, x, y=None, training=False, want_pred=True, x_events=None, y_event=None):
        """x: CUDA fp32 (T,B,3,H,W); y: CUDA fp32 (B,nv,2,nv,nv) or None.
        Returns dict(pred (B,nv,2,nv,nv), loss (0-d CUDA tensor or None), update_gates)."""
        T, B = int(x.shape[0]), int(x.shape[1])
        if B != self.B:
            raise ValueError('batch size is static (lib/layers.py:304-308): built for %d, got %d'
                             % (self.B, B))
        if tuple(x.shape[2:]) != (3, self.img, self.img):
            raise ValueError('expected views of shape (3,%d,%d), got %s' % (self.img, self.img, tuple(x.shape[2:])))
        if training and y is None:
            raise ValueError('training needs ground-truth voxels')
        TB = T * B
        ctx = Ctx(self.dtype, self.algo, self.wgrad_algo, training)
        self.ctx = ctx
        if self.split:
            ops.split_cache_clear()
        if training:
            for p in self.P.values():
                p.grad_written = False
            # bias gradients accumulate (several of them are formed inside data-gradient epilogues):
            # the biases sit together at the end of the flat store, one memset clears them all
            st = self.store
            if st.n > st.n_decay:
                ops.fill_zero(st.g[st.n_decay:])
                for n, p in self.P.items():
                    if n.endswith('.b') or '.b_' in n:
                        p.grad_written = True
        P = self.P
        x = x.contiguous()
        head = None
        if self.rowpack and x_events is not None:
            # x is still arriving from the host, one view per event (Net._stage_inputs): run the first
            # layers view by view so that they overlap the remaining copies (the views are the outermost
            # axis of every channels-last tensor, so a view is a contiguous slice); the tape entries are
            # made below on the full tensors
            cur = torch.cuda.current_stream()
            img, rp = self.img, self.rowpack
            names = ['conv1a', 'conv1b'] if self.residual else ['conv1']
            po = img // 2 + 1
            xin = ctx.empty(TB * img * img * rp, x)
            ys = [ctx.empty(TB * img * img * P[n + '.W'].cout, x) for n in names]
            pooled = ctx.empty(TB * po * po * P[names[-1] + '.W'].cout, x)
            pcode = (torch.empty(pooled.numel(), dtype=torch.uint8, device=x.device)
                     if (training and ops.POOL_CODES) else None)
            for t in range(T):
                cur.wait_event(x_events[t])
                sl = lambda buf, per: buf[t * B * per:(t + 1) * B * per]
                ops.nchw_to_rowpack(x[t], sl(xin, img * img * rp), 7, rp)
                src, cin = sl(xin, img * img * rp), rp
                for i, n in enumerate(names):
                    W_, b_ = P[n + '.W'], P[n + '.b']
                    kk = (1, 7, 1) if i == 0 else (1, 3, 3)
                    ops.useful_frac = 21.0 / rp if i == 0 else 1.0
                    dst = sl(ys[i], img * img * W_.cout)
                    ops.conv_fwd_x(src, W_.operand(), b_.data, None, None, dst, (B, 1, img, img, cin), W_.cout, kk,
                                 L.EPI_BIAS | L.EPI_LRELU, ctx.algo)
                    ops.useful_frac = 1.0
                    src, cin = dst, W_.cout
                ops.maxpool_fwd(src, None, sl(pooled, po * po * cin), B, img, img, cin,
                                None if pcode is None else sl(pcode, po * po * cin))
            head = (xin, ys, pooled, pcode)
        elif x_events is not None:
            # host inputs staged on a side stream (Net._stage_inputs) but no chunked head on this path
            # (fp32 FFMA kernels): every view must have landed before the first kernel reads x
            cur = torch.cuda.current_stream()
            for ev in x_events:
                cur.wait_event(ev)
        if head is not None:
            a = Act(head[0], (TB, 1, self.img, self.img, self.rowpack))
            a.useful = 21.0 / self.rowpack
        elif self.rowpack:
            xin = ctx.empty(TB * self.img * self.img * self.rowpack, x)
            ops.nchw_to_rowpack(x.view(TB, 3, self.img, self.img), xin, 7, self.rowpack)
            a = Act(xin, (TB, 1, self.img, self.img, self.rowpack))
            a.useful = 21.0 / self.rowpack         # 7 taps x 3 channels of the 32 are real
        else:
            xin = ctx.empty(TB * self.img * self.img * self.cin_pad, x)
            ops.nchw_to_nhwc(x.view(TB, 3, self.img, self.img), xin, self.cin_pad)
            a = Act(xin, (TB, 1, self.img, self.img, self.cin_pad))
            a.useful = 3.0 / self.cin_pad          # RGB padded with zero channels

        # row-packed first layer: packed channel 21 is the constant 1 -> bias gradient = dW[:, centre tap, 21]
        ones = (3, 21) if (self.rowpack and self.rowpack > 21) else None

        def c2(name, inp, lrelu, residual=None):
            k = 7 if name in ('conv1a', 'conv1') else (1 if name.endswith('c') else 3)
            kk = (1, k, 1) if (k == 7 and self.rowpack) else (1, k, k)
            return ctx.conv(inp, P[name + '.W'], P[name + '.b'], kk, lrelu, residual,
                            bias_via_ones=ones if (k == 7 and self.rowpack) else None)

        def c3(name, inp, lrelu, residual=None, out_dtype=None, up2=False):
            k = 1 if name == 'conv9c' else 3
            return ctx.conv(inp, P[name + '.W'], P[name + '.b'], (k, k, k), lrelu, residual, out_dtype, up2)

        # ---- encoder over all T*B views (hoisted out of the scan) ----------------------------
        def c2h(name, inp, i):          # chunked head: outputs exist already
            kk = (1, 7, 1) if i == 0 else (1, 3, 3)
            return ctx.conv(inp, P[name + '.W'], P[name + '.b'], kk, True, out_data=head[1][i],
                            bias_via_ones=ones if i == 0 else None)

        if self.residual:
            if head is not None:
                pool1 = ctx.pool(c2h('conv1b', c2h('conv1a', a, 0), 1), out_data=head[2], out_code=head[3])
            else:
                pool1 = ctx.pool(c2('conv1b', c2('conv1a', a, True), True))
            if self.debug is not None:
                self.debug['xin'], self.debug['pool1'] = a.data, pool1.data
            mark_mid = len(ctx.tape)            # tape entries from here on: conv2a ... (enc_mid, enc_hi)
            pool2 = ctx.pool(c2('conv2c', pool1, False), c2('conv2b', c2('conv2a', pool1, True), True))
            pool3 = ctx.pool(c2('conv3c', pool2, False), c2('conv3b', c2('conv3a', pool2, True), True))
            mark_hi = len(ctx.tape)             # from here on: conv4a ... fc7 (enc_hi)
            pool4 = ctx.pool(c2('conv4b', c2('conv4a', pool3, True), True))
            pool5 = ctx.pool(c2('conv5c', pool4, False), c2('conv5b', c2('conv5a', pool4, True), True))
            pool6 = ctx.pool(pool5, c2('conv6b', c2('conv6a', pool5, True), True))
        else:
            t_ = a
            for i in range(6):       # conv -> pool -> lrelu == conv -> lrelu -> pool (monotone)
                if i == 1:
                    mark_mid = len(ctx.tape)
                if i == 3:
                    mark_hi = len(ctx.tape)
                if i == 0 and head is not None:
                    t_ = ctx.pool(c2h('conv1', t_, 0), out_data=head[2], out_code=head[3])
                else:
                    t_ = ctx.pool(c2('conv%d' % (i + 1), t_, True))
            pool6 = t_
        flat = ctx.reshape(pool6, (TB, 1, 1, 1, N_CONVFILTER[5] * 9))
        f = ctx.conv(flat, P['fc7.W'], P['fc7.b'], (1, 1, 1), True)
        if self.debug is not None:
            self.debug['pool6'], self.debug['f'] = pool6.data, f.data
        mark_enc = len(ctx.tape)
        self._enc_marks = {'enc_hi': mark_hi, 'enc_mid': mark_mid}

        # ---- recurrence ---------------------------------------------------------------------
        if self.lstm:
            h, gates = self.lstm_sequence(ctx, f, T, B)
        else:
            h, gates = self.gru_sequence(ctx, f, T, B)

        mark_rnn = len(ctx.tape)
        self._marks = (mark_enc, mark_rnn)

        # ---- decoder ------------------------------------------------------------------------
        if self.residual and ctx.zero_skip(HID_C, HID_C):
            # Unpool3D is never materialised for the convolutions (UP2: 1/8 of the MACs, the data
            # gradient lands on the small tensor); only the residual adds at 8^3 / 16^3 see a
            # zero-stuffed tensor (res_gru_net.py:171-178)
            rect7 = c3('conv7b', c3('conv7a', h, True, up2=True), True)
            res7 = ctx.add_act(ctx.unpool(h), rect7)
            rect8 = c3('conv8b', c3('conv8a', res7, True, up2=True), True)
            res8 = ctx.add_act(ctx.unpool(res7), rect8)
            rect9 = c3('conv9b', c3('conv9a', res8, True, up2=True), True)
            res9 = c3('conv9c', res8, False, residual=rect9, up2=True)
            rect10a = c3('conv10a', res9, True)
            rect10 = c3('conv10b', rect10a, True)
            res10 = c3('conv10c', rect10a, False, residual=rect10)
            logits = c3('conv11', res10, False, out_dtype=torch.float32)
        elif self.residual:
            unpool7 = ctx.unpool(h)
            rect7 = c3('conv7b', c3('conv7a', unpool7, True), True)
            unpool8 = ctx.unpool(unpool7, rect7)                      # unpool(res7)
            rect8 = c3('conv8b', c3('conv8a', unpool8, True), True)
            unpool9 = ctx.unpool(unpool8, rect8)                      # unpool(res8)
            rect9 = c3('conv9b', c3('conv9a', unpool9, True), True)
            res9 = c3('conv9c', unpool9, False, residual=rect9)
            rect10a = c3('conv10a', res9, True)
            rect10 = c3('conv10b', rect10a, True)
            res10 = c3('conv10c', rect10a, False, residual=rect10)
            logits = c3('conv11', res10, False, out_dtype=torch.float32)
        else:
            t_ = c3('conv7', h, True, up2=True)       # falls back to unpool + conv off the tensor path
            t_ = c3('conv8', t_, True, up2=True)
            t_ = c3('conv9', t_, True, up2=True)
            t_ = c3('conv10', t_, True)
            logits = c3('conv11', t_, False, out_dtype=torch.float32)

        # ---- SoftmaxWithLoss3D --------------------------------------------------------------
        nv = self.n_vox
        pred = torch.empty(B, nv, 2, nv, nv, dtype=torch.float32, device=x.device) if want_pred else None
        loss_sum = torch.zeros(1, dtype=torch.float32, device=x.device) if y is not None else None
        dlogits = None
        if training:
            dlogits = ctx.empty(B * nv ** 3 * self.logit_c, x)
        if y_event is not None:
            torch.cuda.current_stream().wait_event(y_event)
        ops.softmax_ce3d(logits.data, None if y is None else y.contiguous(), pred, loss_sum, dlogits,
                         B, nv, 1.0 / (B * nv ** 3), self.logit_c)
        if training:
            logits.grad, logits.n_recv = dlogits, 1
        if self.debug is not None:
            self.debug['h'], self.debug['logits'] = h.data, logits.data
        loss = None if loss_sum is None else loss_sum[0] / float(B * nv ** 3)
        g5 = N_GRU_VOX
        gates_ref = gates.reshape(T, B, g5, g5, g5, HID_C).permute(0, 1, 2, 5, 3, 4)
        self.last = dict(pred=pred, loss=loss, update_gates=gates_ref, logits=logits.data)
        return self.last

    def forward_graph(self, x, y=None):
        """Inference through a CUDA graph: the ~60-120 launches of a forward pass (each a ctypes call, ~10 us
        of host time) are captured once per (views, has-y) and replayed as ONE launch -- at batch 1 the
        pass is host-launch bound (0.9 ms), not GPU bound.  Outputs live in the graph's static buffers and
        are valid until the next replay."""
        T = int(x.shape[0])
        key = (T, y is not None)
        ent = self._graphs.get(key)        # (weights are read from the same buffers at replay: no invalidation)
        if ent is None:
            sx = torch.empty_like(x)
            sy = None if y is None else torch.empty_like(y)
            sx.copy_(x)
            if y is not None:
                sy.copy_(y)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):          # warm-up: lazy allocations, function attributes, err word
                for _ in range(2):
                    self.forward(sx, sy, training=False)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            # thread-local capture mode: the input pipeline's staging thread (lib/data_process.DeviceBatchQueue)
            # allocates pinned / device memory and copies on its own stream while e.g. a validation pass is being
            # captured here; under the default global mode any such call from another thread invalidates the capture
            import gc
            gc_was = gc.isenabled()
            if os.environ.get('R2N2_GRAPH_GC_OFF', '1') != '0':
                gc.disable()            # a cyclic-GC pass inside the capture could destroy old graphs / pools (cudaFree)
            try:
                with torch.cuda.graph(g, capture_error_mode='thread_local'):
                    out = self.forward(sx, sy, training=False)
            finally:
                if gc_was:
                    gc.enable()
            self.ctx = None
            ent = dict(graph=g, x=sx, y=sy, out=out, version=self.store.version)
            self._graphs[key] = ent
        ent['x'].copy_(x)
        if y is not None:
            ent['y'].copy_(y)
        ent['graph'].replay()
        self.last = ent['out']
        return ent['out']

    def backward(self):
        """Reverse pass over the tape: fills the flat gradient buffer (store.g)."""
        ctx = self.ctx
        if ctx is None or not ctx.training:
            raise RuntimeError('backward() needs a preceding forward(training=True)')
        mark_enc, mark_rnn = self._marks
        tape = ctx.tape
        for i in range(len(tape) - 1, -1, -1):
            tape[i]()
            if i == mark_rnn:
                self._finalize_stage('decoder')
            if i == mark_enc:
                self._finalize_stage('rnn')
            if i == self._enc_marks.get('enc_hi'):
                self._finalize_stage('enc_hi')
            if i == self._enc_marks.get('enc_mid'):
                self._finalize_stage('enc_mid')
        ctx.tape = []
        self._finalize_stage('enc_lo')
        self.ctx = None
        if self.split:
            ops.split_cache_clear()

    def stage_segments(self):
        """Non-bias store segments per backward stage, in the order their gradients become final:
        decoder (conv7..11) -> recurrent unit -> enc_hi (fc7, conv6..conv4) -> enc_mid (conv3, conv2) ->
        enc_lo (conv1a/conv1b: the only weights whose all-reduce nothing is left to hide behind)."""
        out = {'decoder': [], 'rnn': [], 'enc_hi': [], 'enc_mid': [], 'enc_lo': []}
        for name in self.P:
            if name.endswith('.b') or '.b_' in name:
                continue
            if name.startswith('rnn.'):
                out['rnn'].append(name)
            elif name.startswith('fc7'):
                out['enc_hi'].append(name)
            else:
                n = int(''.join(ch for ch in name.split('.')[0][4:] if ch.isdigit()))
                out['decoder' if n >= 7 else ('enc_hi' if n >= 4 else ('enc_mid' if n >= 2 else 'enc_lo'))].append(name)
        return out

    def _finalize_stage(self, stage):
        names = self.stage_segments()[stage]
        if stage == 'enc_lo':       # biases ride with the last stage
            names = names + [n for n in self.P if n.endswith('.b') or '.b_' in n]
        for n in names:
            p = self.P[n]
            if not p.grad_written:
                p.grad.zero_()
                p.grad_written = True
        if self.progress is not None:
            self.progress(stage)

"""COMPUTE='bf16x3': the fp32-accuracy path on the tensor cores (verdict r1, missing #2).  Every contraction is three
tcgen05 passes over hi/lo bf16 operand halves (r2n2_split_bf16), accumulated in fp32 (R2N2_EPI_RES_PRE).  The reference
computes in float32 (.theanorc:2); north-star bound: voxel probabilities within 1e-3 max-abs, IoU@0.4 identical to 3
decimals."""
import numpy as np
import pytest
import torch

import r2n2_b200  # noqa: F401
from r2n2_b200 import ops, _lib as L
from r2n2_b200.lib.config import cfg
from r2n2_b200.lib.solver import Solver
from r2n2_b200.models import load_model
from oracle import ref_torch as rt
import emu

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _r(n, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, generator=g) * scale


def test_split_bf16_kernel():
    for n in (8, 13, 4096 + 5, 1 << 20):
        src = _r(n, 1) * torch.logspace(-6, 2, n)
        hi = torch.empty(n, dtype=torch.bfloat16, device=DEV)
        lo = torch.empty(n, dtype=torch.bfloat16, device=DEV)
        ops.split_bf16(src.to(DEV), hi, lo)
        wh, wl = torch.empty(n, dtype=torch.bfloat16), torch.empty(n, dtype=torch.bfloat16)
        emu.split_bf16(src, wh, wl)
        assert torch.equal(hi.cpu(), wh) and torch.equal(lo.cpu(), wl)
        rec = hi.float().cpu().double() + lo.float().cpu().double()
        assert ((rec - src.double()).abs() <= src.abs().double() * 2.0 ** -16).all()


# per-tap kernel, flat-halo kernel (plain / dx-stacked), UP2 / DOWN2 geometry, FC
RES_PRE_CASES = [
    (2, 1, 16, 16, 64, 128, (1, 3, 3), 0),
    (3, 1, 17, 17, 96, 96, (1, 3, 3), 0),          # conv1b shape class (flat-halo when large enough)
    (2, 1, 130, 130, 96, 96, (1, 3, 3), 0),        # flat-halo kernel for sure
    (1, 32, 32, 32, 32, 32, (3, 3, 3), 0),         # flat-halo, resident filter
    (1, 32, 32, 32, 64, 64, (3, 3, 3), 0),         # flat-halo dx-stacked
    (5, 1, 1, 1, 2304, 1024, (1, 1, 1), 0),
    (2, 4, 4, 4, 128, 128, (3, 3, 3), L.CONV_UP2),
    (2, 4, 4, 4, 128, 64, (3, 3, 3), L.CONV_DOWN2),
]


@pytest.mark.parametrize('case', RES_PRE_CASES)
@pytest.mark.parametrize('epi', [0, L.EPI_BIAS | L.EPI_LRELU, L.EPI_BIAS | L.EPI_MASK])
def test_conv_fwd_res_pre(case, epi):
    """R2N2_EPI_RESIDUAL | R2N2_EPI_RES_PRE: the fp32 residual enters BEFORE bias / LeakyReLU / mask."""
    N, D, H, W, Cin, Cout, k, geo = case
    taps = k[0] * k[1] * k[2]
    xs = 8 if geo & L.CONV_DOWN2 else 1
    ys = 8 if geo & L.CONV_UP2 else 1
    P = N * D * H * W
    bf = torch.bfloat16
    x = _r(P * xs * Cin, seed=1).to(bf)
    w = _r(Cout * taps * Cin, seed=2, scale=(2.0 / (taps * Cin)) ** 0.5).to(bf)
    b = _r(Cout, seed=3)
    res = _r(P * ys * Cout, seed=4)
    mask = _r(P * ys * Cout, seed=5)
    flags = geo | epi | L.EPI_RESIDUAL | L.EPI_RES_PRE
    want = torch.empty(P * ys * Cout, dtype=torch.float32)
    emu.conv_fwd(x.float(), w.float(), b, mask, res, want, (N, D, H, W, Cin), Cout, k, flags, 0)
    y = res.to(DEV).clone()                                   # in place: y is its own residual
    ops.conv_fwd(x.to(DEV), w.to(DEV), b.to(DEV), mask.to(DEV), y, y, (N, D, H, W, Cin), Cout, k, flags, L.ALGO_UMMA)
    torch.cuda.synchronize()
    assert ops.umma_error_flag() == 0
    err = (y.cpu() - want).abs().max().item()
    assert err <= 2e-3 * max(1.0, want.abs().max().item()), (case, epi, err)


@pytest.mark.parametrize('case', [(2, 1, 16, 16, 64, 128, (1, 3, 3)), (1, 16, 16, 16, 32, 32, (3, 3, 3)),
                                  (7, 1, 1, 1, 1024, 512, (1, 1, 1)), (2, 1, 33, 33, 96, 128, (1, 3, 3))])
def test_split_contraction_reaches_fp32_accuracy(case):
    """conv_fwd_x / conv_wgrad_x with ALGO_SPLIT on fp32 data vs the exact (float64) contraction: ~1e-5 relative,
    where one bf16 pass is ~4e-3."""
    N, D, H, W, Cin, Cout, k = case
    taps = k[0] * k[1] * k[2]
    P = N * D * H * W
    x = _r(P * Cin, seed=1)
    w = _r(Cout * taps * Cin, seed=2, scale=(2.0 / (taps * Cin)) ** 0.5)
    b = _r(Cout, seed=3)
    dy = _r(P * Cout, seed=6)
    want = torch.empty(P * Cout, dtype=torch.float64)
    emu.conv_fwd(x.double(), w.double(), b.double(), None, None, want, (N, D, H, W, Cin), Cout, k, L.EPI_BIAS | L.EPI_LRELU, 0)
    ops.split_cache_clear()
    wd = w.to(DEV)
    whi = torch.empty_like(wd, dtype=torch.bfloat16)
    wlo = torch.empty_like(wd, dtype=torch.bfloat16)
    ops.split_bf16(wd, whi, wlo)
    y = torch.empty(P * Cout, dtype=torch.float32, device=DEV)
    xd = x.to(DEV)
    ops.conv_fwd_x(xd, ops.SplitPair(whi, wlo, wd), b.to(DEV), None, None, y, (N, D, H, W, Cin), Cout, k,
                   L.EPI_BIAS | L.EPI_LRELU, L.ALGO_SPLIT)
    e3 = ((y.cpu().double() - want).abs().max() / want.abs().max()).item()
    y1 = torch.empty(P * Cout, dtype=torch.float32, device=DEV)
    ops.conv_fwd(xd.to(torch.bfloat16), whi, b.to(DEV), None, None, y1, (N, D, H, W, Cin), Cout, k,
                 L.EPI_BIAS | L.EPI_LRELU, L.ALGO_UMMA)
    e1 = ((y1.cpu().double() - want).abs().max() / want.abs().max()).item()
    wantw = torch.zeros(Cout * taps * Cin, dtype=torch.float64)
    emu.conv_wgrad(x.double(), dy.double(), wantw, (N, D, H, W, Cin), Cout, k, 0, False)
    dw = torch.empty(Cout * taps * Cin, dtype=torch.float32, device=DEV)
    ops.conv_wgrad_x(xd, dy.to(DEV), dw, (N, D, H, W, Cin), Cout, k, L.ALGO_SPLIT, False)
    torch.cuda.synchronize()
    ew = ((dw.cpu().double() - wantw).abs().max() / wantw.abs().max()).item()
    ops.split_cache_clear()
    print('%s: fwd rel err bf16x3 %.2e (one bf16 pass %.2e), wgrad %.2e' % (case, e3, e1, ew))
    assert ops.umma_error_flag() == 0
    assert e3 < 3e-5 and ew < 3e-5 and e1 > 10 * e3


def _build(net_name, B, compute, params=None, seed=0):
    cfg.CONST.BATCH_SIZE = B
    cfg.CONST.COMPUTE = compute
    try:
        net = load_model(net_name)()
    finally:
        cfg.CONST.COMPUTE = 'fp32'
    if params is None:
        params = rt.init_params(net_name, seed=seed)
    net.engine.set_params(params)
    return net, params


@pytest.mark.parametrize('net_name,B,T', [('ResidualGRUNet', 2, 3), ('GRUNet', 2, 2), ('ResidualLSTMNet', 2, 2)])
def test_forward_parity_bf16x3(net_name, B, T):
    net, params = _build(net_name, B, 'bf16x3')
    x, y = rt.synthetic_batch(T, B)
    pred, loss, acts = Solver(net).test_output(x, y)
    with torch.no_grad():
        out = rt.RefNet(net_name, params).forward(x, y)
    rp = out['pred'].numpy()
    err = float(np.abs(pred - rp).max())
    print('%s bf16x3: max|dp| %.3e loss %.6f vs %.6f gates %.3e' % (net_name, err, loss, float(out['loss']),
                                                                  np.abs(acts[0] - out['update_gates'].numpy()).max()))
    assert err < 2e-4                          # north-star bound 1e-3; measured ~2e-5
    assert abs(loss - float(out['loss'])) < 2e-5 * max(1.0, abs(float(out['loss'])))
    assert round(rt.iou(pred, y, 0.4), 3) == round(rt.iou(rp, y, 0.4), 3)
    assert ops.umma_error_flag() == 0


def test_gradients_and_adam_bf16x3():
    """tensor.grad parity vs the float64 oracle at the level real float32 reaches (tests/test_nets_gpu.py), and
    3 Adam steps (hi / lo operand mirrors refreshed by the extra split pass) against the oracle's."""
    net, params = _build('ResidualGRUNet', 2, 'bf16x3')
    x, y = rt.synthetic_batch(2, 2)
    out = net.forward_backward(x, y)
    grads = net.get_grads()
    ref = rt.RefNet('ResidualGRUNet', params, dtype=torch.float64)
    rout, rgrads = ref.loss_and_grads(x, y)
    assert abs(float(out['loss']) - float(rout['loss'].detach())) < 2e-5
    l2 = lambda a, b: np.linalg.norm((a - b).ravel()) / (np.linalg.norm(b.ravel()) + 1e-30)
    worst = ('', 0.0)
    for (name, _, _), g, rg in zip(ref.spec, grads, rgrads):
        e = l2(g, rg.numpy())
        if e > worst[1]:
            worst = (name, e)
        assert e < 2e-2, (name, e)
    print('bf16x3 grads: worst relative L2 vs fp64 oracle %.3e (%s)' % (worst[1], worst[0]))
    solver = Solver(net)
    solver.set_lr(1e-4)
    r32 = rt.RefNet('ResidualGRUNet', params)
    radam = rt.RefAdam(r32, lr=1e-4)
    for it in range(3):
        l = solver.train_loss(x, y)
        rl, _ = radam.step(x, y)
        # (Adam's sign-like first steps amplify the ~1e-5 gradient differences: measured 1.1e-4 at the third step)
        assert abs(l - rl) < 3e-4 * max(1.0, abs(rl)), (it, l, rl)

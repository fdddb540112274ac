"""GPU parity tests of the whole hot path through the reference-facing API (models + Solver):
forward probabilities, loss, update gates, gradients, Adam steps and IoU@0.4 against the CPU
oracle on the same seeded inputs and weights.

Tolerances (BASELINE.json north_star): fp32 path: voxel probabilities max-abs <= 1e-3 (we assert
1e-4), IoU@0.4 identical to 3 decimals.  bf16/tcgen05 path: stated tolerance 3e-2 max-abs on
probabilities (measured values are printed), IoU within 5e-3.
"""
import numpy as np
import pytest
import torch

import r2n2_b200  # noqa: F401
from r2n2_b200.lib.config import cfg
from r2n2_b200.lib.solver import Solver
from r2n2_b200.models import load_model
from oracle import ref_torch as rt

pytestmark = pytest.mark.gpu


def _build(net_name, B, compute='fp32', seed=0):
    cfg.CONST.BATCH_SIZE = B
    cfg.CONST.COMPUTE = compute
    try:
        net = load_model(net_name)()
    finally:
        cfg.CONST.COMPUTE = 'fp32'
    params = rt.init_params(net_name, seed=seed)
    net.engine.set_params(params)
    return net, params


@pytest.mark.parametrize('net_name,B,T', [('ResidualGRUNet', 2, 3), ('GRUNet', 3, 1), ('GRUNet', 2, 2),
                                          ('ResidualLSTMNet', 2, 2)])
def test_forward_parity_fp32(net_name, B, T):
    net, params = _build(net_name, B)
    x, y = rt.synthetic_batch(T, B)
    solver = Solver(net)
    pred, loss, acts = solver.test_output(x, y)
    ref = rt.RefNet(net_name, params)
    with torch.no_grad():
        out = ref.forward(x, y)
    rp = out['pred'].numpy()
    err = np.abs(pred - rp).max()
    print('%s fp32: max|dp| %.3e, loss %.6f vs %.6f' % (net_name, err, loss, float(out['loss'])))
    assert pred.shape == (B, 32, 2, 32, 32) and acts[0].shape == (T, B, 4, 128, 4, 4)
    assert err < 1e-4
    assert abs(loss - float(out['loss'])) < 1e-5 * max(1.0, abs(float(out['loss'])))
    assert np.abs(acts[0] - out['update_gates'].numpy()).max() < 1e-4
    assert round(rt.iou(pred, y, 0.4), 3) == round(rt.iou(rp, y, 0.4), 3)
    # y=None path: (prediction, activations), zeros fed for y (lib/solver.py:222-238)
    p2, a2 = solver.test_output(x)
    assert np.array_equal(p2, pred)


def test_demo_shape_config1():
    """BASELINE config 1: ResidualGRUNet, batch 1, 3 views (demo.py path, random-init weights)."""
    net, params = _build('ResidualGRUNet', 1)
    x, _ = rt.synthetic_batch(3, 1, seed_x=7)
    pred, acts = Solver(net).test_output(x)
    with torch.no_grad():
        out = rt.RefNet('ResidualGRUNet', params).forward(x)
    assert np.abs(pred - out['pred'].numpy()).max() < 1e-4
    occ_ref = out['pred'].numpy()[0, :, 1] > 0.4          # demo.py:69
    assert np.array_equal(pred[0, :, 1] > 0.4, occ_ref) or \
        np.mean((pred[0, :, 1] > 0.4) != occ_ref) < 1e-4


@pytest.mark.parametrize('net_name,B,T', [('ResidualGRUNet', 2, 2), ('GRUNet', 2, 3), ('ResidualLSTMNet', 1, 2)])
def test_gradients_fp32(net_name, B, T):
    """tensor.grad parity.  The yardstick is the oracle itself: its float32 run differs from its
    float64 run by fp32 round-off amplified through max-pool arg-max / LeakyReLU sign decisions;
    the CUDA fp32 path must be as close to the float64 oracle as that (x3), or within 1e-3."""
    net, params = _build(net_name, B)
    x, y = rt.synthetic_batch(T, B)
    out = net.forward_backward(x, y)
    grads = net.get_grads()
    ref = rt.RefNet(net_name, params, dtype=torch.float64)
    rout, rgrads = ref.loss_and_grads(x, y)
    ref32 = rt.RefNet(net_name, params, dtype=torch.float32)
    _, rgrads32 = ref32.loss_and_grads(x, y)
    assert abs(float(out['loss']) - float(rout['loss'].detach())) < 1e-5
    worst = 0.0
    l2 = lambda a, b: np.linalg.norm((a - b).ravel()) / (np.linalg.norm(b.ravel()) + 1e-30)
    for (name, _, _), g, rg, rg32 in zip(ref.spec, grads, rgrads, rgrads32):
        rg = rg.numpy()
        e_gpu, e_o32 = l2(g, rg), l2(rg32.numpy().astype(np.float64), rg)
        worst = max(worst, e_gpu)
        # measured (tools/diag_grads.py): both the CUDA path and the fp32 oracle sit at 1e-7 for
        # conv11 and grow smoothly to ~2-4e-3 at conv1a (run-to-run variable on the oracle side)
        assert e_gpu < max(6e-3, 3 * e_o32), (name, e_gpu, e_o32)
        assert np.abs(g - rg).max() / (np.abs(rg).max() + 1e-30) < 2e-2, name
    print('%s grads: worst relative L2 error vs fp64 oracle %.3e' % (net_name, worst))


def test_adam_training_steps_fp32():
    """3 Solver.train_loss steps == 3 oracle Adam steps (lib/solver.py:20-48, 102-109)."""
    net, params = _build('ResidualGRUNet', 2)
    x, y = rt.synthetic_batch(2, 2)
    solver = Solver(net)
    solver.set_lr(1e-4)
    ref = rt.RefNet('ResidualGRUNet', params)
    radam = rt.RefAdam(ref, lr=1e-4)
    for it in range(3):
        l = solver.train_loss(x, y)
        rl, _ = radam.step(x, y)
        assert abs(l - rl) < 2e-5 * max(1.0, abs(rl)), (it, l, rl)
    assert float(solver.iteration) == 3.0
    # 3 steps at lr 1e-4 move a parameter by at most 3e-4.  Adam's first steps are sign-like (m / sqrt(v) = +-1):
    # where a gradient element is ~0 in fp32, rounding decides the direction, so single elements may differ by a
    # sizeable part of the distance travelled.  The informative bounds: the bulk (99 %) agrees to 5e-5 (1/6 of the
    # movement), at most 0.2 % (or 2) of any tensor's elements are further than 1e-4 apart, none further than 2 x movement.
    for (name, _, _), a, b, p0 in zip(ref.spec, net.engine.get_params(), ref.P, params):
        d = np.abs(a - b.detach().numpy())
        assert np.quantile(d, 0.99) < 5e-5, (name, np.quantile(d, 0.99))
        assert (d > 1e-4).sum() <= max(2, 2e-3 * d.size), (name, int((d > 1e-4).sum()), d.size)
        assert d.max() <= 6.1e-4, (name, d.max())
    # save / load round trip in the reference's weights.npy format
    import os, tempfile
    with tempfile.TemporaryDirectory() as td:
        solver.save([1.0, 0.5], td, 7)
        assert os.path.islink(os.path.join(td, 'weights.npy'))
        net2, _ = _build('ResidualGRUNet', 2, seed=5)
        net2.load(os.path.join(td, 'weights.npy'))
        for a, b in zip(net.engine.get_params(), net2.engine.get_params()):
            assert np.array_equal(a, b)


@pytest.mark.parametrize('compute', ['fp32', 'bf16'])
def test_optimiser_pass_in_ranges_equals_one_pass(compute):
    """Solver._apply_updates(ranges=...) (the optional data-parallel split update, dist.GradSync.wait_early): updating the
    flat store range by range leaves parameters, moments and the bf16 operand mirror bit-identical to the single launch,
    wherever the cuts fall relative to the L2-decay boundary.  (Same gradient buffer for both: the weight-gradient kernels
    accumulate with atomics, two backward passes differ in the last bits.)"""
    net, _ = _build('GRUNet', 2, compute)
    solver = Solver(net)
    solver.set_lr(1e-3)
    x, y = rt.synthetic_batch(2, 2)
    st = net.engine.store
    n = st.p.numel()
    cuts = [(0, 64 * 1000), (64 * 1000, st.n_decay - 64 * 3), (st.n_decay - 64 * 3, st.n_decay + 64), (st.n_decay + 64, n)]
    names = ['p', 'm', 'v'] + (['mirror'] if st.mirror is not None else [])
    for it in range(2):
        solver.iteration = np.float32(solver.iteration + 1)
        net.forward_backward(x, y)
        st.ensure_moments()
        before = {k: getattr(st, k).clone() for k in names}
        solver._apply_updates(1.0)
        one = {k: getattr(st, k).clone() for k in names}
        for k in names:
            getattr(st, k).copy_(before[k])
        solver._apply_updates(1.0, ranges=cuts[2:], finish=False)       # late ranges first: order must not matter
        solver._apply_updates(1.0, ranges=cuts[:2])
        for k in names:
            assert torch.equal(one[k], getattr(st, k)), (it, k)
        assert not torch.equal(one['p'], before['p'])


def test_load_weights_file_written_the_reference_way(tmp_path):
    """f-2: a checkpoint written exactly as /root/reference/models/net.py:60-67 writes it --
    `np.save(filename, [param.val.get_value() for param in self.params])`, i.e. a pickled OBJECT array of ragged
    float32 arrays -- loads, and an .npz written the way models/net.py:73-76 reads one does too."""
    net, params = _build('ResidualGRUNet', 1, seed=1)
    want = rt.init_params('ResidualGRUNet', seed=9)
    fn = str(tmp_path / 'weights.npy')
    arr = np.empty(len(want), dtype=object)          # numpy >= 1.24 refuses the implicit ragged conversion
    arr[:] = [np.asarray(w) for w in want]
    np.save(fn, arr)
    net.load(fn)
    for a, b in zip(net.engine.get_params(), want):
        assert np.array_equal(a, b)
    # short file + ignore_param (models/net.py:78-86)
    np.save(fn, arr[:10])
    net.load(fn, ignore_param=True)
    with pytest.raises(IndexError):
        net.load(fn, ignore_param=False)
    fz = str(tmp_path / 'weights.npz')
    np.savez(fz, arr)
    net2, _ = _build('ResidualGRUNet', 1, seed=2)
    net2.load(fz)
    for a, b in zip(net2.engine.get_params(), want):
        assert np.array_equal(a, b)
    # the forward pass of the loaded weights is the oracle's
    x, _ = rt.synthetic_batch(1, 1)
    pred, _ = Solver(net2).test_output(x)
    with torch.no_grad():
        out = rt.RefNet('ResidualGRUNet', want).forward(x)
    assert np.abs(pred - out['pred'].numpy()).max() < 1e-4


@pytest.mark.parametrize('compute', ['fp32', 'bf16'])
def test_user_defined_net_trains_on_gpu(compute):
    """VERDICT r1 missing #3: a network_definition written against the layer classes (here the reference's
    ResidualGRUNet body, models/layerwise.py) trains through Net.forward_backward / Solver.train_loss.
    fp32: gradients == the fused engine's (which are oracle-pinned); bf16: within the bf16 tolerance."""
    from r2n2_b200.models import layerwise
    from r2n2_b200.lib import layers as ly
    B, T = 2, 2
    net, params = _build('ResidualGRUNet', B)
    x, y = rt.synthetic_batch(T, B)
    oe = net.forward_backward(x, y)
    ge = net.get_grads()
    cfg.CONST.BATCH_SIZE, cfg.CONST.COMPUTE = B, compute
    n0 = len(ly.trainable_params)
    try:
        lw = layerwise.LayerwiseResidualGRUNet()
        for w, v in zip(lw._weights, params):
            w.val.set_value(v)
        ol = lw.forward_backward(x, y)
        gl = lw.get_grads()
        tol_loss, tol_g = (1e-5, 2e-3) if compute == 'fp32' else (5e-3, 0.9)      # measured 4.3e-4 / 0.47 (B=2)
        assert abs(float(ol['loss']) - float(oe['loss'])) < tol_loss * max(1.0, abs(float(oe['loss'])))
        worst = 0.0
        for a, b in zip(gl, ge):
            worst = max(worst, float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)))
        print('layerwise %s: worst per-tensor gradient rel-L2 vs engine fp32 %.3e' % (compute, worst))
        assert worst < tol_g
        s = Solver(lw)
        s.set_lr(1e-4)
        losses = [s.train_loss(x, y) for _ in range(4)]
        assert np.isfinite(losses).all() and losses[-1] < losses[0]
        if compute == 'fp32':
            se = Solver(net)
            se.set_lr(1e-4)
            le = [se.train_loss(x, y) for _ in range(4)]
            # (measured 2.5e-5 after 4 steps: sign-like first Adam steps amplify the 4e-4 gradient difference)
            assert np.abs(np.array(le) - np.array(losses)).max() < 1e-4
        pred, loss, acts = s.test_output(x, y)
        assert pred.shape == (B, 32, 2, 32, 32) and acts[0].shape == (T, B, 4, 128, 4, 4)
    finally:
        cfg.CONST.COMPUTE = 'fp32'
        del ly.trainable_params[n0:]


def test_layer_api_path_matches_engine():
    """reference-ordered, layer-by-layer graph (drop-in layer classes) == fused engine."""
    net, params = _build('ResidualGRUNet', 1)
    x, y = rt.synthetic_batch(2, 1)
    out_e = net.forward(x, y)
    out_l = net.forward_layerwise(x, y)
    assert (out_e['pred'] - out_l['pred']).abs().max().item() < 1e-5
    assert abs(float(out_e['loss']) - float(out_l['loss'])) < 1e-6
    assert (out_e['update_gates'].float() - out_l['update_gates']).abs().max().item() < 1e-5


def test_static_shapes_and_errors():
    net, _ = _build('GRUNet', 2)
    x, y = rt.synthetic_batch(1, 3)
    with pytest.raises(ValueError):       # batch size is static in the graph (quirk 14)
        Solver(net).test_output(x)
    with pytest.raises(ValueError):
        net.engine.set_params([np.zeros(3)])


@pytest.mark.parametrize('net_name,B,T', [('ResidualGRUNet', 2, 3), ('GRUNet', 2, 2)])
def test_forward_parity_bf16_umma(net_name, B, T):
    net, params = _build(net_name, B, compute='bf16')
    x, y = rt.synthetic_batch(T, B)
    pred, loss, acts = Solver(net).test_output(x, y)
    with torch.no_grad():
        out = rt.RefNet(net_name, params).forward(x, y)
    rp = out['pred'].numpy()
    err = np.abs(pred - rp).max()
    iou_a, iou_b = rt.iou(pred, y, 0.4), rt.iou(rp, y, 0.4)
    print('%s bf16/tcgen05: max|dp| %.3e mean|dp| %.3e loss %.5f vs %.5f IoU %.4f vs %.4f'
          % (net_name, err, np.abs(pred - rp).mean(), loss, float(out['loss']), iou_a, iou_b))
    assert err < 3e-2                      # stated tolerance of the bf16 tensor-core path
    assert abs(loss - float(out['loss'])) < 2e-2 * max(1.0, abs(float(out['loss'])))
    assert abs(iou_a - iou_b) < 5e-3


def test_gradients_bf16_umma():
    net, params = _build('ResidualGRUNet', 2, compute='bf16')
    x, y = rt.synthetic_batch(2, 2)
    out = net.forward_backward(x, y)
    grads = net.get_grads()
    ref = rt.RefNet('ResidualGRUNet', params)
    rout, rgrads = ref.loss_and_grads(x, y)
    worst = 0.0
    for (name, _, _), g, rg in zip(ref.spec, grads, rgrads):
        rg = rg.numpy()
        # bf16 operands: compare direction and scale of each gradient tensor
        cos = float((g * rg).sum() / (np.linalg.norm(g) * np.linalg.norm(rg) + 1e-30))
        worst = min(worst, cos) if worst else cos
        assert cos > 0.9, (name, cos)
    print('bf16 grads: worst cosine %.4f' % worst)


def test_long_recurrence_config5():
    """BASELINE config 5 (24-view sequence inference, long recurrence of the fused GRU kernels), reduced
    batch: fp32 path against the oracle at T=8, then the bf16/tcgen05 path against our own fp32 path at
    T=24 (the oracle's 24 CPU encoders would dominate the test time)."""
    net, params = _build('ResidualGRUNet', 1)
    x, y = rt.synthetic_batch(8, 1, seed_x=11)
    pred, loss, acts = Solver(net).test_output(x, y)
    ref = rt.RefNet('ResidualGRUNet', params)
    with torch.no_grad():
        out = ref.forward(x, y)
    err = np.abs(pred - out['pred'].numpy()).max()
    print('T=8 fp32: max|dp| %.3e' % err)
    assert err < 1e-4 and acts[0].shape == (8, 1, 4, 128, 4, 4)
    x24, y24 = rt.synthetic_batch(24, 1, seed_x=12)
    p32, _, a32 = Solver(net).test_output(x24, y24)
    net16, _ = _build('ResidualGRUNet', 1, compute='bf16')
    p16, _, a16 = Solver(net16).test_output(x24, y24)
    e16 = np.abs(p16 - p32).max()
    print('T=24 bf16 vs fp32: max|dp| %.3e, update gates %.3e' % (e16, np.abs(a16[0] - a32[0]).max()))
    assert np.isfinite(p16).all() and e16 < 3e-2
    assert abs(rt.iou(p16, y24, 0.4) - rt.iou(p32, y24, 0.4)) < 5e-3


def test_evaluation_loop_matches_host_metrics(tmp_path):
    """lib/test_net.py mirror: per-batch cost, per-sample IoU counts (lib/voxel.py:4-11) and average precision
    (sklearn.metrics.average_precision_score) computed on the GPU == the host restatements on the same
    predictions; result.mat has the reference's keys."""
    from oracle import ref_numpy as rn
    from r2n2_b200.lib import test_net as tn
    import scipy.io as sio
    net, params = _build('GRUNet', 2)
    solver = Solver(net)
    batches = [rt.synthetic_batch(2, 2, seed_x=21), rt.synthetic_batch(1, 2, seed_x=22)]
    cfg.DIR.OUT_PATH = str(tmp_path)
    old = list(cfg.TEST.VOXEL_THRESH)
    cfg.TEST.VOXEL_THRESH = [0.4, 0.5]
    try:
        res = tn.test_net(solver, batches)
    finally:
        cfg.TEST.VOXEL_THRESH = old
    assert res['cost'].shape == (2,) and res['mAP'].shape == (2, 2) and res['0.4'].shape == (2, 2, 5)
    for bi, (x, y) in enumerate(batches):
        pred, loss, _ = solver.test_output(x, y)
        assert abs(res['cost'][bi] - loss) < 1e-6
        for j in range(2):
            want_ap = rn.average_precision(y[j, :, 1], pred[j, :, 1])
            assert abs(res['mAP'][bi, j] - want_ap) <= 1e-12
            for th in (0.4, 0.5):
                assert np.array_equal(res[str(th)][bi, j], np.asarray(rn.evaluate_voxel_prediction(pred[j], y[j], th), np.float64))
    mat = sio.loadmat(str(tmp_path / 'test' / 'result.mat'))
    assert {'cost', 'mAP', '0.4', '0.5'} <= set(mat.keys())


def test_demo_script_writes_reference_obj(tmp_path):
    """demo.py end to end: PNG views -> Net.load(weights.npy) -> Solver.test_output -> voxel2obj.  The OBJ
    must be the text the reference's writer produces for the oracle's thresholded prediction."""
    import hashlib
    import importlib.util
    import os
    from PIL import Image
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('r2n2_demo', os.path.join(root, 'demo.py'))
    demo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(demo)
    r = np.random.RandomState(5)
    paths = []
    for i in range(3):
        a = (r.rand(127, 127, 4) * 255).astype(np.uint8)
        paths.append(str(tmp_path / ('%d.png' % i)))
        Image.fromarray(a, 'RGBA').save(paths[-1])
    net, params = _build('ResidualGRUNet', 1)
    wfile = str(tmp_path / 'weights.npy')
    net.save(wfile)
    out_obj = str(tmp_path / 'prediction.obj')
    assert demo.main([out_obj, '--weights', wfile, '--images'] + paths) == 0
    x = demo.load_demo_images(paths)
    assert x.shape == (3, 1, 3, 127, 127) and x.dtype == np.float32
    ref = rt.RefNet('ResidualGRUNet', params)
    with torch.no_grad():
        p = ref.forward(x, None)['pred'].numpy()
    from r2n2_b200.lib.voxel import voxel2obj
    ref_obj = str(tmp_path / 'ref.obj')
    voxel2obj(ref_obj, p[0, :, 1] > 0.4)
    margin = np.abs(p[0, :, 1] - 0.4).min()
    if margin > 1e-4:                                   # no voxel sits on the threshold: identical meshes
        assert hashlib.sha256(open(out_obj, 'rb').read()).hexdigest() == \
            hashlib.sha256(open(ref_obj, 'rb').read()).hexdigest()
    assert open(out_obj).readline() == 'g\n'
    assert demo.main([out_obj, '--weights', str(tmp_path / 'missing.npy'), '--images'] + paths) == 1

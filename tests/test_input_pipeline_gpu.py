"""Input pipeline on the GPU: r2n2_voxel_labels, the raw-batch conversion and
DeviceBatchQueue against batches produced by the REAL reference workers
(tests/golden/make_data_process_fixture.py); bit-exact."""
import multiprocessing
import os
import queue
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import r2n2_b200  # noqa: F401,E402
from r2n2_b200 import ops  # noqa: E402
from r2n2_b200.lib import data_io, data_process  # noqa: E402
from r2n2_b200.lib.config import cfg  # noqa: E402
from r2n2_b200.lib.solver import Solver  # noqa: E402
from r2n2_b200.models import load_model  # noqa: E402
import dataset_synth  # noqa: E402
from test_data_process import RUNS, golden, raw_batches, restore_cfg  # noqa: F401,E402

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def no_leftover_pipeline_threads():
    """Every DeviceBatchQueue must be closed by its user: a staging thread that outlives its test would keep
    allocating / copying behind later tests (and CUDA-graph captures)."""
    yield
    import threading
    left = [t.name for t in threading.enumerate() if t.name.startswith('r2n2-device-batch-queue') and t.is_alive()]
    assert not left, 'staging threads still alive: %s' % left


def test_voxel_labels_kernel():
    r = np.random.RandomState(3)
    for B, nv in ((1, 32), (5, 32), (3, 7)):
        occ = r.randint(0, 3, (B, nv, nv, nv)).astype(np.uint8)
        occ[0, 0, 0, 0] = 255
        y = ops.voxel_labels(torch.from_numpy(occ).cuda()).cpu().numpy()
        assert y.shape == (B, nv, 2, nv, nv)
        assert np.array_equal(y[:, :, 0], (occ == 0).astype(np.float32))
        assert np.array_equal(y[:, :, 1], (occ == 1).astype(np.float32))


@pytest.mark.parametrize('tag', sorted(RUNS))
def test_raw_batches_to_device_match_reference(tag, tmp_path, restore_cfg):  # noqa: F811
    _, items = raw_batches(tag, tmp_path)
    for i, (views, occ) in enumerate(items):
        img, vox = golden(tag, i)
        x, y = data_process.batch_to_device(views, occ)
        assert x.is_cuda and x.dtype == torch.float32 and tuple(x.shape) == img.shape
        assert np.array_equal(x.cpu().numpy(), img)
        assert np.array_equal(y.cpu().numpy(), vox)


def test_device_batch_queue_order_and_passthrough(tmp_path, restore_cfg):  # noqa: F811
    _, items = raw_batches('train', tmp_path)
    q = queue.Queue()
    for it in items * 3:                                   # 9 batches through 3 pinned slots: slots are recycled
        q.put(it)
    img0, vox0 = golden('train', 0)
    q.put((img0, vox0))                                    # a float batch from a reference-style worker
    dq = data_process.DeviceBatchQueue(q, depth=2)
    try:
        for k in range(9):
            x, y = dq.get(timeout=60)
            img, vox = golden('train', k % 3)
            assert np.array_equal(x.cpu().numpy(), img) and np.array_equal(y.cpu().numpy(), vox)
        x, y = dq.get(timeout=60)
        assert np.array_equal(x.cpu().numpy(), img0) and np.array_equal(y.cpu().numpy(), vox0)
        with pytest.raises(queue.Empty):
            dq.get(block=False)
    finally:
        dq.close()


def test_device_batch_queue_reports_bad_batches():
    q = queue.Queue()
    rgba = np.zeros((1, 2, 130, 130, 4), np.uint8)
    params = np.zeros((1, 2, 6), np.int32)
    params[0, 1, 0] = 9                                    # 9 + 127 > 130: crop window leaves the render
    q.put((data_process.RawViews(rgba, params), np.zeros((2, 32, 32, 32), np.uint8)))
    dq = data_process.DeviceBatchQueue(q)
    try:
        with pytest.raises(RuntimeError):
            dq.get(timeout=30)
    finally:
        dq.close()


def test_solver_train_from_worker_processes(tmp_path, restore_cfg):  # noqa: F811
    """Workers -> mp.Queue -> DeviceBatchQueue -> Solver.train (lib/train_net.py:55-85 wiring): the losses equal
    those of the same batches fed as host float arrays."""
    dataset_synth.apply_cfg(cfg, dataset_synth.build(str(tmp_path)), batch_size=2)
    pairs = data_io.category_model_id_pair(dataset_portion=[0, 1])
    keep = (cfg.TRAIN.NUM_ITERATION, cfg.DIR.OUT_PATH, cfg.TRAIN.SAVE_FREQ)
    cfg.TRAIN.NUM_ITERATION, cfg.DIR.OUT_PATH, cfg.TRAIN.SAVE_FREQ = 2, str(tmp_path / 'out'), 10 ** 6
    try:
        np.random.seed(21)
        ref_proc = data_process.ReconstructionDataProcess(None, pairs, repeat=True, train=True)
        host_batches = []
        for _ in range(3):
            views, occ = ref_proc.make_batch()
            x, y = data_process.batch_to_device(views, occ)
            host_batches.append((x.cpu().numpy(), y.cpu().numpy()))

        def run(train_queue):
            torch.manual_seed(0)
            np.random.seed(0)
            net = load_model('GRUNet')()
            return Solver(net).train(train_queue)

        hq = queue.Queue()
        for b in host_batches:
            hq.put(b)
        losses_host = run(hq)

        np.random.seed(21)                                   # the forked worker inherits this RNG state
        mq = multiprocessing.Queue(2)
        procs = data_process.make_data_processes(mq, pairs, 1, repeat=True, train=True)
        dq = data_process.DeviceBatchQueue(mq, depth=2)
        try:
            losses_dev = run(dq)
        finally:
            dq.close()
            data_process.kill_processes(mq, procs)
        assert len(losses_host) == 3 and len(losses_dev) == 3
        # same inputs and weights; the host path runs the first layers view by view behind the copies and the
        # loss / weight-gradient sums are accumulated atomically, so equality holds to rounding only
        assert np.allclose(losses_host, losses_dev, rtol=1e-4)
    finally:
        cfg.TRAIN.NUM_ITERATION, cfg.DIR.OUT_PATH, cfg.TRAIN.SAVE_FREQ = keep


def test_train_net_and_test_net_entry_points(tmp_path, restore_cfg):  # noqa: F811
    """lib/train_net.train_net() and lib/test_net.test_net() called as main.py calls them (main.py:95-108):
    everything comes from cfg."""
    import scipy.io as sio
    from r2n2_b200.lib import test_net as tn, train_net as trn
    dataset_synth.apply_cfg(cfg, dataset_synth.build(str(tmp_path / 'data')), batch_size=2)
    keep = {k: cfg.TRAIN[k] for k in ('NUM_ITERATION', 'SAVE_FREQ', 'VALIDATION_FREQ', 'NUM_VALIDATION_ITERATIONS',
                                      'DATASET_PORTION')}
    keep_c = (cfg.CONST.NETWORK_CLASS, cfg.CONST.WEIGHTS, cfg.DIR.OUT_PATH, cfg.TEST.DATASET_PORTION)
    try:
        cfg.CONST.NETWORK_CLASS = 'GRUNet'
        cfg.DIR.OUT_PATH = str(tmp_path / 'out')
        cfg.TRAIN.NUM_ITERATION, cfg.TRAIN.SAVE_FREQ = 2, 2
        cfg.TRAIN.VALIDATION_FREQ, cfg.TRAIN.NUM_VALIDATION_ITERATIONS = 2, 1
        cfg.TRAIN.DATASET_PORTION, cfg.TEST.DATASET_PORTION = [0, 0.67], [0, 1]
        np.random.seed(5)
        losses = trn.train_net()
        assert len(losses) == 3 and all(np.isfinite(losses))
        assert trn.train_processes is None                          # workers were stopped
        wfile = os.path.join(cfg.DIR.OUT_PATH, 'weights.npy')      # symlink written by Solver.save at step 2
        assert os.path.exists(wfile)
        cfg.CONST.WEIGHTS = wfile
        res = tn.test_net()
        mat = sio.loadmat(os.path.join(cfg.DIR.OUT_PATH, cfg.TEST.EXP_NAME, 'result.mat'))
        assert res['cost'].shape == (3,) and res['mAP'].shape == (3, 2) and res['0.4'].shape == (3, 2, 5)
        assert np.allclose(mat['mAP'], res['mAP']) and np.isfinite(res['cost']).all()
    finally:
        for k, v in keep.items():
            cfg.TRAIN[k] = v
        cfg.CONST.NETWORK_CLASS, cfg.CONST.WEIGHTS, cfg.DIR.OUT_PATH, cfg.TEST.DATASET_PORTION = keep_c


def test_graph_capture_while_pipeline_thread_runs(tmp_path, restore_cfg):  # noqa: F811
    """Solver.train validates through Solver.test_output, whose first call per view count captures a CUDA graph,
    while the training queue's staging thread keeps allocating and copying: the capture must survive it."""
    _, items = raw_batches('train', tmp_path)
    q = queue.Queue()
    for it in items * 40:
        q.put(it)
    dq = data_process.DeviceBatchQueue(q, depth=2)
    try:
        cfg.CONST.BATCH_SIZE = 4
        net = load_model('GRUNet')()
        solver = Solver(net)
        for k in range(6):
            x, y = dq.get(timeout=60)                       # keeps the thread busy between captures
            T = 1 + k % 3                                   # three different view counts: three captures
            pred, loss, _ = solver.test_output(x[:T].contiguous(), y)
            assert np.isfinite(pred).all() and np.isfinite(loss)
    finally:
        dq.close()

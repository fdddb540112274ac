"""Input-pipeline workers (lib/data_process.py mirror) against batches produced by the REAL reference workers
on the same synthetic dataset and numpy seed (tests/golden/make_data_process_fixture.py).  The raw batches are
turned into float batches here with the oracle's preprocess_img (CPU); tests/test_input_pipeline_gpu.py does
the same through the kernels."""
import multiprocessing
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import dataset_synth  # noqa: E402
from oracle import ref_numpy as rn  # noqa: E402
from r2n2_b200.lib import data_io, data_process  # noqa: E402
from r2n2_b200.lib.config import cfg  # noqa: E402

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'data_process.npz'))
RUNS = {'train': (11, True, True, [0, 1], None), 'test': (12, False, False, [0, 1], None),
        'rgb': (13, True, True, [0, 0.67], 1)}


class ListQueue(object):
    def __init__(self):
        self.items = []

    def put(self, item, block=True):
        self.items.append(item)


@pytest.fixture
def restore_cfg():
    keep = (cfg.DATASET, dict(cfg.DIR), cfg.CONST.BATCH_SIZE, cfg.TRAIN.NUM_RENDERING)
    yield
    cfg.DATASET, cfg.CONST.BATCH_SIZE, cfg.TRAIN.NUM_RENDERING = keep[0], keep[2], keep[3]
    cfg.DIR.update(keep[1])


def raw_batches(tag, root):
    """The mirror's workers run in-process on the fixture's dataset / seed; returns (pairs, [(RawViews, occ)])."""
    seed, train, repeat, portion, rgb = RUNS[tag]
    dataset_synth.apply_cfg(cfg, dataset_synth.build(str(root), rgb_model=rgb))
    np.random.seed(seed)
    pairs = data_io.category_model_id_pair(dataset_portion=portion)
    proc = data_process.ReconstructionDataProcess(ListQueue(), pairs, repeat=repeat, train=train)
    n = int(G['n_' + tag])
    if repeat:
        items = [proc.make_batch() for _ in range(n)]
    else:
        proc.run()                                   # stops by itself after the short last batch
        items = proc.data_queue.items
    return pairs, items


def golden(tag, i):
    img = G['img_%s_%d' % (tag, i)].astype(np.float32) / np.float32(255)
    B = img.shape[1]
    v1 = np.unpackbits(G['vox_%s_%d' % (tag, i)])[:B * 32 ** 3].reshape(B, 32, 32, 32)
    v0 = np.unpackbits(G['vox0_%s_%d' % (tag, i)])[:B * 32 ** 3].reshape(B, 32, 32, 32)
    return img, np.stack([v0, v1], axis=2).astype(np.float32)


@pytest.mark.parametrize('tag', sorted(RUNS))
def test_worker_batches_match_reference(tag, tmp_path, restore_cfg):
    pairs, items = raw_batches(tag, tmp_path)
    assert ['/'.join(p) for p in pairs] == G['pairs_' + tag].tolist()
    assert len(items) == int(G['n_' + tag])
    for i, (views, occ) in enumerate(items):
        img, vox = golden(tag, i)
        assert views.shape == img.shape and views.rgba.dtype == np.uint8 and occ.dtype == np.uint8
        T, B = views.rgba.shape[:2]
        got = np.zeros(img.shape, np.float32)
        for t in range(T):
            for b in range(B):
                got[t, b] = rn.preprocess_img(views.rgba[t, b], views.params[t, b]).transpose(2, 0, 1)
        assert np.array_equal(got, img)
        y = np.stack([occ == 0, occ == 1], axis=2).astype(np.float32)
        assert np.array_equal(y, vox)


def test_short_last_batch_has_unfilled_slots(tmp_path, restore_cfg):
    _, items = raw_batches('test', tmp_path)
    views, occ = items[-1]
    assert (occ[2:] == 2).all() and (occ[:2] < 2).all()
    assert (views.rgba[:, 2:] == 0).all() and (views.params[:, 2:] == 0).all()


def test_process_and_queue_protocol(tmp_path, restore_cfg):
    """A real worker process, the mp.Queue, get_while_running and kill_processes (lib/data_process.py:176-216)."""
    dataset_synth.apply_cfg(cfg, dataset_synth.build(str(tmp_path)))
    pairs = data_io.category_model_id_pair(dataset_portion=[0, 1])
    q = multiprocessing.Queue(2)
    procs = data_process.make_data_processes(q, pairs, 1, repeat=False, train=False)
    got = list(data_process.get_while_running(procs[0], q, sleep_time=0.01))
    assert len(got) == 2
    for views, occ in got:
        assert isinstance(views, data_process.RawViews) and views.rgba.shape[1:] == (4, 137, 137, 4)
        assert occ.shape == (4, 32, 32, 32)
    q2 = multiprocessing.Queue(2)
    procs = data_process.make_data_processes(q2, pairs, 2, repeat=True, train=True)
    views, occ = q2.get(timeout=60)
    assert views.rgba.shape[1] == 4
    data_process.kill_processes(q2, procs)
    for p in procs:
        p.join(timeout=10)
        assert not p.is_alive()


def test_bad_render_size_is_reported(tmp_path, restore_cfg):
    from PIL import Image
    paths = dataset_synth.build(str(tmp_path))
    dataset_synth.apply_cfg(cfg, paths)
    pairs = data_io.category_model_id_pair(dataset_portion=[0, 1])
    cat, mid = pairs[1]
    for i in range(dataset_synth.NUM_RENDERING):
        Image.fromarray(np.zeros((100, 100, 4), np.uint8), 'RGBA').save(data_io.get_rendering_file(cat, mid, i))
    proc = data_process.ReconstructionDataProcess(ListQueue(), pairs, repeat=False, train=False)
    with pytest.raises(ValueError):
        proc.make_batch()


def test_device_queue_needs_cuda():
    """No CPU fallback: the trainer-side half of the pipeline refuses to run without a CUDA device."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    with pytest.raises(RuntimeError):
        data_process.DeviceBatchQueue(ListQueue())
    views = data_process.RawViews(np.zeros((1, 1, 137, 137, 4), np.uint8), np.zeros((1, 1, 6), np.int32))
    assert views.shape == (1, 1, 3, 127, 127)
    with pytest.raises(RuntimeError):
        data_process.batch_to_device(views, np.zeros((1, 32, 32, 32), np.uint8))

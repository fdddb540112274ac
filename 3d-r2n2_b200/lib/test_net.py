"""Evaluation loop of the reference (lib/test_net.py:18-84) on the GPU-side metrics: for every batch
`solver.test_output(batch_img, batch_voxel)`, then per sample the IoU counts of lib/voxel.py:4-11 at every
threshold of cfg.TEST.VOXEL_THRESH (r2n2_voxel_eval) and the average precision that the reference gets from
sklearn.metrics.average_precision_score (r2n2_average_precision).  Same `result.mat` keys: 'cost', 'mAP' and
one (num_batch, batch_size, 5) array per threshold.

`batches` is any iterable of (batch_img (T,B,3,H,W), batch_voxel (B,n_vox,2,n_vox,n_vox)) float32 arrays or
device tensors, or of the raw batches the lib/data_process.py workers emit, e.g. their queue drained with
get_while_running()."""
import os

import numpy as np
import torch

from .config import cfg
from .data_process import RawViews, batch_to_device
from .. import ops


def evaluate_batches(solver, batches, num_batch=None, thresholds=None, verbose=True):
    thresholds = list(cfg.TEST.VOXEL_THRESH) if thresholds is None else list(thresholds)
    results = {'cost': [], 'mAP': []}
    for t in thresholds:
        results[str(t)] = []
    for batch_idx, (batch_img, batch_voxel) in enumerate(batches):
        if num_batch is not None and batch_idx == num_batch:
            break
        if isinstance(batch_img, RawViews):                  # raw batch of lib/data_process.py workers
            batch_img, batch_voxel = batch_to_device(batch_img, batch_voxel)
        pred, loss, _ = solver.test_output(batch_img, batch_voxel)
        p = torch.as_tensor(pred).cuda().contiguous()
        if not torch.is_tensor(batch_voxel):
            batch_voxel = torch.as_tensor(np.asarray(batch_voxel, dtype=np.float32))
        g = batch_voxel.float().cuda().contiguous()
        for t in thresholds:
            results[str(t)].append(ops.voxel_eval(p, g, float(t)).cpu().numpy().astype(np.float64))
        ap = ops.average_precision(p, g).cpu().numpy()
        results['mAP'].append(ap)
        results['cost'].append(float(loss))
        if verbose:
            print('%d/%s, costs: %f, mAP: %f' % (batch_idx, '?' if num_batch is None else num_batch, loss, ap.mean()))
    out = {k: np.asarray(v) for k, v in results.items()}
    if verbose and len(out['cost']):
        print('Total loss: %f' % np.mean(out['cost']))
        print('Total mAP: %f' % np.mean(out['mAP']))
    return out


def test_net(solver=None, batches=None, num_batch=None):
    """lib/test_net.py:18-84; writes result.mat.  Called without arguments it does what the reference does:
    builds cfg.CONST.NETWORK_CLASS, loads cfg.CONST.WEIGHTS, starts one non-repeating test worker on
    cfg.TEST.DATASET_PORTION and evaluates int(num_data / batch_size) batches."""
    import scipy.io as sio
    procs = queue = None
    if solver is None:
        from ..models import load_model
        from .solver import Solver
        net = load_model(cfg.CONST.NETWORK_CLASS)(compute_grad=False)
        net.load(cfg.CONST.WEIGHTS)
        solver = Solver(net)
    if batches is None:
        from multiprocessing import Queue
        from .data_io import category_model_id_pair
        from .data_process import get_while_running, make_data_processes
        queue = Queue(cfg.QUEUE_SIZE)
        data_pair = category_model_id_pair(dataset_portion=cfg.TEST.DATASET_PORTION)
        procs = make_data_processes(queue, data_pair, 1, repeat=False, train=False)
        num_batch = int(len(data_pair) / cfg.CONST.BATCH_SIZE)
        batches = get_while_running(procs[0], queue)
    result_dir = os.path.join(cfg.DIR.OUT_PATH, cfg.TEST.EXP_NAME)
    if not os.path.exists(result_dir):
        os.makedirs(result_dir)
    result_fn = os.path.join(result_dir, 'result.mat')
    print("Exp file will be written to: " + result_fn)
    try:
        results = evaluate_batches(solver, batches, num_batch)
    finally:
        if procs:
            from .data_process import kill_processes
            kill_processes(queue, procs)
    sio.savemat(result_fn, results)
    return results

"""Evaluation loop of the reference (lib/test_net.py:18-84) on the GPU-side metrics: for every batch
`solver.test_output(batch_img, batch_voxel)`, then per sample the IoU counts of lib/voxel.py:4-11 at every
threshold of cfg.TEST.VOXEL_THRESH (r2n2_voxel_eval) and the average precision that the reference gets from
sklearn.metrics.average_precision_score (r2n2_average_precision).  Same `result.mat` keys: 'cost', 'mAP' and
one (num_batch, batch_size, 5) array per threshold.

Data loading (lib/data_process.py, lib/data_io.py) is outside the hot path: `batches` is any iterable of
(batch_img (T,B,3,H,W), batch_voxel (B,n_vox,2,n_vox,n_vox)) float32 arrays, e.g. the reference's queue
drained with get_while_running()."""
import os

import numpy as np
import torch

from .config import cfg
from .. import ops


def evaluate_batches(solver, batches, num_batch=None, thresholds=None, verbose=True):
    thresholds = list(cfg.TEST.VOXEL_THRESH) if thresholds is None else list(thresholds)
    results = {'cost': [], 'mAP': []}
    for t in thresholds:
        results[str(t)] = []
    for batch_idx, (batch_img, batch_voxel) in enumerate(batches):
        if num_batch is not None and batch_idx == num_batch:
            break
        pred, loss, _ = solver.test_output(batch_img, batch_voxel)
        p = torch.as_tensor(pred).cuda().contiguous()
        g = torch.as_tensor(np.asarray(batch_voxel, dtype=np.float32)).cuda().contiguous()
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


def test_net(solver, batches, num_batch=None):
    """lib/test_net.py:18-84 with the network / solver built by the caller; writes result.mat."""
    import scipy.io as sio
    result_dir = os.path.join(cfg.DIR.OUT_PATH, cfg.TEST.EXP_NAME)
    if not os.path.exists(result_dir):
        os.makedirs(result_dir)
    result_fn = os.path.join(result_dir, 'result.mat')
    print("Exp file will be written to: " + result_fn)
    results = evaluate_batches(solver, batches, num_batch)
    sio.savemat(result_fn, results)
    return results

"""lib/voxel.py:4-11 on the GPU (r2n2_voxel_eval); numpy in / numpy out like the reference."""
import numpy as np
import torch

from .. import ops


def evaluate_voxel_prediction(preds, gt, thresh):
    """preds, gt: (n_vox, 2, n_vox, n_vox) for one sample (the reference's call shape,
    lib/test_net.py:62-66) or (B, n_vox, 2, n_vox, n_vox).  Returns the reference's
    [diff, intersection, union, num_fp, num_fn] (per sample rows if batched)."""
    p = torch.as_tensor(np.asarray(preds, dtype=np.float32))
    g = torch.as_tensor(np.asarray(gt, dtype=np.float32))
    single = p.dim() == 4
    if single:
        p, g = p[None], g[None]
    counts = ops.voxel_eval(p.cuda().contiguous(), g.cuda().contiguous(), thresh).cpu().numpy()
    return counts[0] if single else counts


def average_precision(preds, gt):
    """Per-sample average precision of the occupancy channel, as lib/test_net.py:69-70 computes it with
    sklearn.metrics.average_precision_score(gt[j, :, 1].flatten(), pred[j, :, 1].flatten()).
    preds, gt: (B, n_vox, 2, n_vox, n_vox) (or one sample); returns float64 [B] (or a float)."""
    p = torch.as_tensor(np.asarray(preds, dtype=np.float32))
    g = torch.as_tensor(np.asarray(gt, dtype=np.float32))
    single = p.dim() == 4
    if single:
        p, g = p[None], g[None]
    ap = ops.average_precision(p.cuda().contiguous(), g.cuda().contiguous()).cpu().numpy()
    return float(ap[0]) if single else ap

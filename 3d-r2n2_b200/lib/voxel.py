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


# ---------------------------------------------------------------------------------------------------
# Mesh export (lib/voxel.py:14-58): host-side post-processing of one thresholded prediction, off the
# hot path.  Same vertices / faces / file text as the reference, computed without the per-voxel loop.
_CUBE_VERTS = np.array([[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1],
                        [1, 0, 0], [1, 0, 1], [1, 1, 0], [1, 1, 1]], np.int64)
_CUBE_FACES = np.array([[0, 1, 2], [1, 3, 2], [2, 3, 6], [3, 7, 6], [0, 2, 6], [0, 6, 4],
                        [0, 5, 1], [0, 4, 5], [6, 7, 5], [6, 5, 4], [1, 7, 3], [1, 5, 7]], np.int64) + 1
_MESH_SCALE = 0.01
_CUBE_DIST_SCALE = 1.1


def voxel2mesh(voxels, surface_view):
    """One cube (8 vertices, 12 triangles, 1-based indices) per voxel above 0.3, in C order of the grid
    (lib/voxel.py:14-43).  With `surface_view` only voxels with an exposed face are kept: the reference
    sums the 3x3x3 neighbourhood of the array after setting the occupied cells to 1 and keeps the voxel
    when the sum is below 27; at index 0 its slice `[i-1:i+2]` is empty and at the last index it is
    clipped, so every voxel on the boundary of the grid is kept.  Like the reference, `voxels` is
    modified in place (occupied cells become 1)."""
    occupied = voxels > 0.3
    voxels[occupied] = 1
    keep = occupied
    if surface_view:
        v = np.asarray(voxels, dtype=np.float64)
        full = np.zeros(v.shape, bool)
        if min(v.shape) >= 3:
            s = v[:-2] + v[1:-1] + v[2:]
            s = s[:, :-2] + s[:, 1:-1] + s[:, 2:]
            s = s[:, :, :-2] + s[:, :, 1:-1] + s[:, :, 2:]
            full[1:-1, 1:-1, 1:-1] = ~(s < 27)
        keep = occupied & ~full
    pos = np.argwhere(keep)                                    # C order, as zip(*np.where(...))
    if len(pos) == 0:
        return np.array([]), np.array([])
    verts = _MESH_SCALE * (_CUBE_VERTS[None, :, :] + _CUBE_DIST_SCALE * pos[:, None, :])
    faces = _CUBE_FACES[None, :, :] + 8 * np.arange(len(pos), dtype=np.int64)[:, None, None]
    return verts.reshape(-1, 3), faces.reshape(-1, 3)


def write_obj(filename, verts, faces):
    """lib/voxel.py:46-58: 'g', a vertex-count comment, 'v x y z' lines (%f), a face-count comment and
    'f a b c' lines."""
    verts, faces = np.asarray(verts), np.asarray(faces)
    with open(filename, 'w') as f:
        f.write('g\n# %d vertex\n' % len(verts))
        if len(verts):
            f.write(''.join('v %f %f %f\n' % (a, b, c) for a, b, c in verts.tolist()))
        f.write('# %d faces\n' % len(faces))
        if len(faces):
            f.write(''.join('f %d %d %d\n' % (a, b, c) for a, b, c in faces.tolist()))


def voxel2obj(filename, pred, surface_view=True):
    """lib/voxel.py:61-63 (called by demo.py:75 with `prediction[0, :, 1] > cfg.TEST.VOXEL_THRESH`)."""
    verts, faces = voxel2mesh(pred, surface_view)
    write_obj(filename, verts, faces)

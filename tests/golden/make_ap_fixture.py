"""Golden vectors for the average-precision metric of the evaluation loop (lib/test_net.py:69-70), produced
by the reference's own dependency: sklearn.metrics.average_precision_score (scikit-learn, requirements.txt)
called exactly as the reference calls it.  Run in the build container:  python tests/golden/make_ap_fixture.py
Cases: smooth random scores, heavily quantised scores (many ties), near-perfect and inverted predictions, a
sample with no positive voxel, a sample with a single positive voxel."""
import os

import numpy as np
import sklearn
import sklearn.metrics

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    r = np.random.RandomState(20240607)
    nv, B = 32, 8
    occ = r.rand(B, nv, nv, nv) < 0.1
    occ[6] = False                                   # no positive voxel
    occ[7] = False
    occ[7, 3, 4, 5] = True                           # exactly one positive
    p1 = r.rand(B, nv, nv, nv).astype(np.float32)
    p1[1] = np.round(p1[1] * 7) / 7                  # 8 distinct values: long runs of ties
    p1[2] = np.where(occ[2], 0.5 + 0.5 * p1[2], 0.5 * p1[2])          # almost perfect
    p1[3] = np.where(occ[3], 0.3 * p1[3], 0.3 + 0.7 * p1[3])          # inverted
    p1[4] = (np.round(p1[4] * 3) / 3).astype(np.float32)
    p1[5] = np.clip(0.3 * occ[5] + 0.7 * p1[5], 0, 1).astype(np.float32)
    pred = np.stack([1 - p1, p1], axis=2).astype(np.float32)          # (B, nv, 2, nv, nv)
    gt = np.stack([~occ, occ], axis=2).astype(np.float32)
    import warnings
    ap = []
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for j in range(B):
            ap.append(sklearn.metrics.average_precision_score(gt[j, :, 1].flatten(), pred[j, :, 1].flatten()))
    np.savez_compressed(os.path.join(HERE, 'average_precision.npz'), p1=p1, occ=np.packbits(occ),
                        ap=np.asarray(ap, np.float64), sklearn_version=sklearn.__version__)
    print('sklearn', sklearn.__version__, 'AP:', ap)


if __name__ == '__main__':
    main()

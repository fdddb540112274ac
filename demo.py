"""Mirror of the reference's demo.py (demo.py:37-86): three views of one object -> voxel prediction ->
prediction.obj.  Usage:

    python demo.py [prediction.obj] [--weights output/ResidualGRUNet/default_model/weights.npy]
                   [--images imgs/0.png imgs/1.png imgs/2.png]

The pretrained weight file is downloaded by the reference; there is no network here, so a missing file is an
error unless --random-init is given (the mesh is then meaningless, the path is the same)."""
import argparse
import os
import sys

import numpy as np

from r2n2_b200.lib.config import cfg
from r2n2_b200.lib.solver import Solver
from r2n2_b200.lib.voxel import voxel2obj
from r2n2_b200.models import load_model

DEFAULT_WEIGHTS = 'output/ResidualGRUNet/default_model/weights.npy'


def load_demo_images(paths):
    """demo.py:37-43: (n_views, 1, 3, H, W) float32 in [0, 1]."""
    from PIL import Image
    ims = []
    for p in paths:
        a = np.array(Image.open(p))
        if a.ndim != 3 or a.shape[2] < 3:
            raise ValueError('%s: expected an RGB(A) image' % p)
        ims.append([a[:, :, :3].transpose((2, 0, 1)).astype(np.float32) / 255.])
    return np.array(ims)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('pred_file_name', nargs='?', default='prediction.obj')
    ap.add_argument('--weights', default=DEFAULT_WEIGHTS)
    ap.add_argument('--images', nargs='+', default=['imgs/%d.png' % i for i in range(3)])
    ap.add_argument('--net', default='ResidualGRUNet')
    ap.add_argument('--random-init', action='store_true')
    args = ap.parse_args(argv)

    demo_imgs = load_demo_images(args.images)
    NetClass = load_model(args.net)
    if NetClass is None:
        return 2
    cfg.CONST.BATCH_SIZE = 1
    net = NetClass(compute_grad=False)
    if os.path.isfile(args.weights):
        net.load(args.weights)
    elif not args.random_init:
        print('weights file %s not found (pass --random-init to run the path without it)' % args.weights)
        return 1
    solver = Solver(net)
    voxel_prediction, _ = solver.test_output(demo_imgs)
    voxel2obj(args.pred_file_name, voxel_prediction[0, :, 1, :, :] > cfg.TEST.VOXEL_THRESH)
    print('wrote %s' % args.pred_file_name)
    return 0


if __name__ == '__main__':
    sys.exit(main())

"""TEST INFRASTRUCTURE: a tiny deterministic ShapeNet-shaped dataset on disk (category json, RGBA renders,
.binvox labels) laid out as lib/config.py:12,32-35 expects.  Shared by tests/golden/make_data_process_fixture.py
(which feeds it to the real reference workers) and the input-pipeline tests."""
import json
import os

import numpy as np

CATS = {'chair': '03001627', 'plane': '02691156'}
MODELS_PER_CAT = 3
NUM_RENDERING = 4
SRC = 137


def _render(r):
    """Blocky (compressible) object on a transparent background, a few semi-transparent pixels."""
    blocks = r.randint(0, 256, (18, 18, 4)).astype(np.uint8)
    im = np.kron(blocks, np.ones((8, 8, 1), np.uint8))[:SRC, :SRC].copy()
    yy, xx = np.mgrid[0:SRC, 0:SRC]
    cy, cx = r.randint(40, 97, 2)
    ry, rx = r.randint(15, 50, 2)
    obj = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
    im[..., 3] = np.where(obj, 255, 0)
    im[::17, ::13, 3] = 128
    return im


def build(root, rgb_model=None):
    """Writes the dataset under `root`; returns the dict of cfg values pointing at it.  `rgb_model` = index of
    a model whose renders are saved without alpha channel (the reference then composites nothing)."""
    from PIL import Image
    from r2n2_b200.lib import binvox_rw
    r = np.random.RandomState(2024)
    cats = {}
    k = 0
    for name, cid in sorted(CATS.items()):
        cats[name] = {'id': cid, 'name': name}
        for m in range(MODELS_PER_CAT):
            mid = 'model%02d_%s' % (m, name)
            vdir = os.path.join(root, 'ShapeNetVox32', cid, mid)
            rdir = os.path.join(root, 'ShapeNetRendering', cid, mid, 'rendering')
            os.makedirs(vdir)
            os.makedirs(rdir)
            occ = r.rand(32, 32, 32) < 0.15
            with open(os.path.join(vdir, 'model.binvox'), 'wb') as f:
                binvox_rw.write(binvox_rw.Voxels(occ, [32, 32, 32], [0.0, 0.0, 0.0], 1.0, 'xyz'), f)
            for i in range(NUM_RENDERING):
                im = _render(r)
                if rgb_model is not None and k == rgb_model:
                    Image.fromarray(im[..., :3].copy(), 'RGB').save(os.path.join(rdir, '%02d.png' % i))
                else:
                    Image.fromarray(im, 'RGBA').save(os.path.join(rdir, '%02d.png' % i))
            k += 1
    ds = os.path.join(root, 'dataset.json')
    with open(ds, 'w') as f:
        json.dump(cats, f)
    return {'DATASET': ds,
            'SHAPENET_QUERY_PATH': os.path.join(root, 'ShapeNetVox32') + '/',
            'VOXEL_PATH': os.path.join(root, 'ShapeNetVox32', '%s', '%s', 'model.binvox'),
            'RENDERING_PATH': os.path.join(root, 'ShapeNetRendering', '%s', '%s', 'rendering')}


def apply_cfg(cfg, paths, batch_size=4):
    cfg.DATASET = paths['DATASET']
    cfg.DIR.SHAPENET_QUERY_PATH = paths['SHAPENET_QUERY_PATH']
    cfg.DIR.VOXEL_PATH = paths['VOXEL_PATH']
    cfg.DIR.RENDERING_PATH = paths['RENDERING_PATH']
    cfg.CONST.BATCH_SIZE = batch_size
    cfg.TRAIN.NUM_RENDERING = NUM_RENDERING

"""Golden vectors for the input pipeline, produced by the REAL reference code:
lib/data_augmentation.preprocess_img (add_random_color_background + image_transform / crop_center + /255.,
lib/data_augmentation.py:6-70) imported from /root/reference, run on synthetic 137x137 RGBA renders with a
seeded numpy RNG.  Run in the build container:  python tests/golden/make_preprocess_fixture.py

Environment shims only (no reference logic changed): `np.float` (removed in numpy 1.24, used at
lib/data_augmentation.py:48) and the `easydict` / `theano` shims of make_reference_fixtures.py.

The random draws are replayed with the same seed to record the parameters the reference used, in ITS order:
randint x3 for the background colour (r, g, b), then (train) randint(0, PAD_Y) = crop row, randint(0, PAD_X) =
crop column, rand() > 0.5 = horizontal flip."""
import os
import sys

import numpy as np
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_reference_fixtures as mrf  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def synth_render(r, size=137):
    """A ShapeNet-like render: transparent background (alpha 0), an opaque object blob, a few stray
    semi-transparent pixels (alpha in 1..254 counts as object: only alpha == 0 takes the background)."""
    yy, xx = np.mgrid[0:size, 0:size]
    cy, cx = r.randint(40, 97, 2)
    ry, rx = r.randint(15, 50, 2)
    obj = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
    im = r.randint(0, 256, (size, size, 4)).astype(np.uint8)
    im[..., 3] = np.where(obj, 255, 0)
    stray = r.rand(size, size) < 0.01
    im[..., 3] = np.where(stray, r.randint(1, 255, (size, size)), im[..., 3])
    return im


def main():
    mrf.install_shims()
    if not hasattr(np, 'float'):
        np.float = float
    sys.path.insert(0, mrf.REF)
    import lib.data_augmentation as ref_aug
    from lib.config import cfg
    assert (cfg.TRAIN.PAD_X, cfg.TRAIN.PAD_Y, cfg.CONST.IMG_W, cfg.CONST.IMG_H) == (10, 10, 127, 127)
    r = np.random.RandomState(99)
    rgba, outs, params, train_flags, seeds = [], [], [], [], []
    for i in range(8):
        im = synth_render(r)
        train = i < 6
        seed = 1000 + i
        np.random.seed(seed)
        out = ref_aug.preprocess_img(Image.fromarray(im, 'RGBA'), train=train)       # (127,127,3)
        # replay the draws
        np.random.seed(seed)
        rng_range = cfg.TRAIN.NO_BG_COLOR_RANGE if train else cfg.TEST.NO_BG_COLOR_RANGE
        bg = [np.random.randint(rng_range[c][0], rng_range[c][1] + 1) for c in range(3)]
        if train:
            cr, cc = np.random.randint(0, cfg.TRAIN.PAD_Y), np.random.randint(0, cfg.TRAIN.PAD_X)
            flip = int(np.random.rand() > 0.5)
        else:
            cr = cc = (137 - 127) // 2
            flip = 0
        rgba.append(im); outs.append(np.asarray(out, np.float32)); params.append([cr, cc, flip] + bg)
        train_flags.append(train); seeds.append(seed)
        assert out.shape == (127, 127, 3) and out.dtype == np.float32, (out.shape, out.dtype)
    np.savez_compressed(os.path.join(HERE, 'preprocess.npz'), rgba=np.stack(rgba), out=np.stack(outs),
                        params=np.asarray(params, np.int32), train=np.asarray(train_flags), seeds=np.asarray(seeds),
                        bg_train=np.asarray(cfg.TRAIN.NO_BG_COLOR_RANGE), bg_test=np.asarray(cfg.TEST.NO_BG_COLOR_RANGE))
    print('params (crop_row, crop_col, flip, r, g, b):')
    print(np.asarray(params))


if __name__ == '__main__':
    main()

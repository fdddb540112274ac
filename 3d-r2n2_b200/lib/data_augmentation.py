"""GPU-side input pipeline: the image part of lib/data_augmentation.py:6-70 / lib/data_process.py:133-148.

The reference decodes every render with PIL and composites / crops / flips / scales it in numpy on the host,
one image at a time, then ships float32 batches through an mp.Queue.  Here the host only draws the random
parameters -- with numpy's global RNG and in the reference's order, so a seeded run makes the same choices --
and ONE kernel (r2n2_preprocess_views) turns the uint8 RGBA renders into the (T, B, 3, H, W) float batch."""
import numpy as np
import torch

from .config import cfg
from .. import ops


def draw_params(n_img, train=True, src_hw=(137, 137)):
    """Per image, in the order of preprocess_img: background colour r, g, b (add_random_color_background,
    lib/data_augmentation.py:41), then for training crop row / column (14) and the flip coin (24); for
    testing the centre crop (31-38) and no flip.  Returns int32 [n_img, 6]."""
    rng_range = cfg.TRAIN.NO_BG_COLOR_RANGE if train else cfg.TEST.NO_BG_COLOR_RANGE
    out = np.zeros((n_img, 6), np.int32)
    for i in range(n_img):
        bg = [np.random.randint(rng_range[c][0], rng_range[c][1] + 1) for c in range(3)]
        if train:
            cr = cc = 0
            if cfg.TRAIN.RANDOM_CROP:
                cr, cc = np.random.randint(0, cfg.TRAIN.PAD_Y), np.random.randint(0, cfg.TRAIN.PAD_X)
            flip = int(cfg.TRAIN.FLIP and np.random.rand() > 0.5)
        else:
            cr, cc = (src_hw[0] - cfg.CONST.IMG_H) // 2, (src_hw[1] - cfg.CONST.IMG_W) // 2
            flip = 0
        out[i] = [cr, cc, flip] + bg
    return out


def preprocess_views(rgba, params=None, train=True, out_dtype=torch.float32):
    """rgba: uint8 array / tensor [n, Hs, Ws, 4] (n = views x batch, any order); returns a CUDA tensor
    [n, 3, IMG_H, IMG_W] in [0, 1].  `params` defaults to draw_params(n, train)."""
    t = torch.as_tensor(rgba)
    if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[3] != 4:
        raise ValueError('expected uint8 RGBA images [n, H, W, 4], got %s %s' % (t.dtype, tuple(t.shape)))
    n, Hs, Ws, _ = t.shape
    if params is None:
        params = draw_params(n, train, (Hs, Ws))
    p = torch.as_tensor(np.asarray(params, np.int32))
    H, W = cfg.CONST.IMG_H, cfg.CONST.IMG_W
    if (p[:, 0].min() < 0 or p[:, 1].min() < 0 or int(p[:, 0].max()) + H > Hs or int(p[:, 1].max()) + W > Ws):
        raise ValueError('crop window outside the source image')
    out = torch.empty(n, 3, H, W, dtype=out_dtype, device='cuda')
    return ops.preprocess_views(t.cuda().contiguous(), p.cuda().contiguous(), out)


def preprocess_img(im, train=True):
    """Reference signature (lib/data_augmentation.py:57-68): one PIL image / HxWx4 uint8 array -> (H, W, 3)
    float32 numpy array."""
    a = np.asarray(im)
    if a.ndim != 3 or a.shape[2] != 4:
        raise ValueError('expected an RGBA image')
    out = preprocess_views(a[None].astype(np.uint8), train=train)
    return out[0].permute(1, 2, 0).contiguous().cpu().numpy()

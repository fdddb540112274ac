"""Golden batches for the input pipeline, produced by the REAL reference workers:
lib/data_process.ReconstructionDataProcess.run (lib/data_process.py:100-173) + lib/data_io.category_model_id_pair
imported from /root/reference and run in-process on the synthetic dataset of tests/dataset_synth.py with a
seeded numpy RNG.  Run in the build container:  python tests/golden/make_data_process_fixture.py

Environment shims only: the `easydict` / inert `theano` shims of make_reference_fixtures.py (the worker reads
theano.config.floatX), `np.float` / `np.bool` / `np.int` (removed from numpy).  The queue is a list that stops
the worker after N batches.

Runs: 'train' = seed 11, training mode, repeat (random view count, crops, flips, colours; a reshuffle);
'test' = seed 12, test mode, no repeat (centre crop, fixed colour, short last batch with zero rows);
'rgb' = seed 13, one model's renders saved without alpha.
Images are stored as round(x * 255) uint8 -- the script asserts float32(k) / 255 reproduces every value."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_reference_fixtures as mrf  # noqa: E402
import dataset_synth  # noqa: E402


class ListQueue(object):
    def __init__(self, proc_ref, limit):
        self.items, self.proc_ref, self.limit = [], proc_ref, limit

    def put(self, item, block=True):
        self.items.append(item)
        if len(self.items) >= self.limit:
            self.proc_ref[0].exit.set()


def run(ref_dp, ref_io, cfg, seed, train, repeat, portion, limit):
    np.random.seed(seed)
    pairs = ref_io.category_model_id_pair(dataset_portion=portion)
    holder = [None]
    q = ListQueue(holder, limit)
    proc = ref_dp.ReconstructionDataProcess(q, pairs, repeat=repeat, train=train)
    holder[0] = proc
    proc.run.__wrapped__(proc) if hasattr(proc.run, '__wrapped__') else proc.run()
    return pairs, q.items


def main():
    mrf.install_shims()
    for n, t in (('float', float), ('bool', bool), ('int', int)):
        if not hasattr(np, n):
            setattr(np, n, t)
    import r2n2_b200  # noqa: F401  (dataset_synth writes the labels with the mirror's writer, itself pinned byte-exact)
    sys.path.insert(0, mrf.REF)
    import lib.data_process as ref_dp
    import lib.data_io as ref_io
    from lib.config import cfg
    out = {}
    for tag, seed, train, repeat, portion, limit, rgb in (('train', 11, True, True, [0, 1], 3, None),
                                                          ('test', 12, False, False, [0, 1], 10, None),
                                                          ('rgb', 13, True, True, [0, 0.67], 1, 1)):
        with tempfile.TemporaryDirectory() as root:
            paths = dataset_synth.build(root, rgb_model=rgb)
            dataset_synth.apply_cfg(cfg, paths)
            pairs, items = run(ref_dp, ref_io, cfg, seed, train, repeat, portion, limit)
        out['pairs_' + tag] = np.array(['/'.join(p) for p in pairs])
        out['n_' + tag] = np.array(len(items))
        for i, (img, vox) in enumerate(items):
            assert img.dtype == np.float32 and vox.dtype == np.float32
            k = np.rint(img * 255).astype(np.uint8)
            assert np.array_equal(k.astype(np.float32) / np.float32(255), img)
            out['img_%s_%d' % (tag, i)] = k
            assert set(np.unique(vox)) <= {0.0, 1.0}
            out['vox_%s_%d' % (tag, i)] = np.packbits(vox[:, :, 1].astype(bool))
            out['vox0_%s_%d' % (tag, i)] = np.packbits(vox[:, :, 0].astype(bool))   # all zero in unfilled rows
            print(tag, i, img.shape, vox.shape, float(img.mean()))
    np.savez_compressed(os.path.join(HERE, 'data_process.npz'), **out)


if __name__ == '__main__':
    main()

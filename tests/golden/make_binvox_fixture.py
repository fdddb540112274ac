"""Golden vectors for the .binvox label format (lib/binvox_rw.py), produced by the REAL reference module
imported from /root/reference.  Shims: `np.bool` / `np.int` (removed from numpy; the module predates that),
its writer emits `chr()` strings, so it is given a latin-1 text file, which stores exactly those bytes, and
0/1 uint8 data (`chr` rejects numpy bools).
Run in the build container:  python tests/golden/make_binvox_fixture.py
Cases: a random 32^3 grid, a sparse 32^3 grid with runs longer than 255, an all-empty and an all-full 8^3
grid, a run of exactly 255 cells, each as the reference writes it (file bytes) and as it reads it back (fix_coords on/off, coordinates)."""
import importlib.util
import io
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    np.bool = bool
    np.int = int
    spec = importlib.util.spec_from_file_location('ref_binvox_rw', '/root/reference/lib/binvox_rw.py')
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    r = np.random.RandomState(99)
    grids = {'rand32': r.rand(32, 32, 32) < 0.3, 'sparse32': r.rand(32, 32, 32) < 0.002,
             'empty8': np.zeros((8, 8, 8), bool), 'full8': np.ones((8, 8, 8), bool)}
    grids['sparse32'][10:20, 5:25, 8:9] = True
    stored = np.zeros(512, bool)                      # a run of exactly 255 followed by a change: the reference
    stored[255:300] = True                            # then emits a zero-length pair
    grids['run255'] = np.transpose(stored.reshape(8, 8, 8), (0, 2, 1))
    out = {}
    for name, g in grids.items():
        d = g.shape[0]
        vox = ref.Voxels(g.astype(np.uint8), [d, d, d], [0.0, -0.25, 1.5], 0.75, 'xyz')
        buf = io.TextIOWrapper(io.BytesIO(), encoding='latin-1', newline='\n')
        ref.write(vox, buf)
        buf.flush()
        data = buf.buffer.getvalue()
        out['file_' + name] = np.frombuffer(data, np.uint8)
        a = ref.read_as_3d_array(io.BytesIO(data))
        b = ref.read_as_3d_array(io.BytesIO(data), fix_coords=False)
        assert np.array_equal(a.data, g)
        out['xyz_' + name] = np.packbits(a.data)
        out['xzy_' + name] = np.packbits(b.data)
        out['meta_' + name] = np.array(a.dims + a.translate + [a.scale], np.float64)
        print(name, len(data), 'bytes', int(g.sum()), 'occupied')
    np.savez_compressed(os.path.join(HERE, 'binvox.npz'), **out)


if __name__ == '__main__':
    main()

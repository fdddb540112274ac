"""Golden vectors for the mesh export (lib/voxel.py:14-63), produced by the REAL reference module imported
from /root/reference (pure numpy, imports as is).  Run in the build container:
    python tests/golden/make_voxel_mesh_fixture.py
Cases: a boolean 32^3 blob with solid interior (what demo.py passes), a float 12^3 grid whose unoccupied
cells hold fractional values (they enter the reference's neighbourhood sum), a full 5^3 block (only the
centre 3^3 is hidden), an empty grid; each with surface_view on and off.  Small cases keep the arrays,
every case keeps the sha256 of the OBJ text."""
import hashlib
import importlib.util
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    r = np.random.RandomState(77)
    g = np.indices((32, 32, 32)).astype(np.float64)
    blob = ((g[0] - 15) ** 2 / 150 + (g[1] - 14) ** 2 / 90 + (g[2] - 17) ** 2 / 200) < 1
    blob |= r.rand(32, 32, 32) < 0.02
    blob[:3, 10:20, 10:20] = True                    # touches the boundary
    fl = r.rand(12, 12, 12)
    fl[3:9, 3:9, 3:9] = 0.9                          # 6^3 block, interior hidden only where neighbours are 1
    return {'blob32': blob, 'float12': fl, 'full5': np.ones((5, 5, 5), bool), 'empty4': np.zeros((4, 4, 4), bool)}


def main():
    spec = importlib.util.spec_from_file_location('ref_voxel', '/root/reference/lib/voxel.py')
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    for name, vox in cases().items():
        out['in_' + name] = np.packbits(vox) if vox.dtype == bool else vox
        for sv in (True, False):
            key = '%s_sv%d' % (name, sv)
            v, f = ref.voxel2mesh(vox.copy(), sv)
            with tempfile.NamedTemporaryFile('r', suffix='.obj') as tmp:
                ref.voxel2obj(tmp.name, vox.copy(), sv)
                text = open(tmp.name).read()
            out['sha_' + key] = hashlib.sha256(text.encode()).hexdigest()
            out['n_' + key] = np.array([len(v), len(f)])
            if vox.size <= 12 ** 3:
                out['verts_' + key], out['faces_' + key] = v, f
            print(key, len(v), len(f), out['sha_' + key][:12])
    np.savez_compressed(os.path.join(HERE, 'voxel_mesh.npz'), **out)


if __name__ == '__main__':
    main()

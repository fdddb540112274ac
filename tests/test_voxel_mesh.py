"""Mesh export (lib/voxel.py:14-63) against golden vectors made by the real reference module
(tests/golden/make_voxel_mesh_fixture.py)."""
import hashlib
import os

import numpy as np
import pytest

from r2n2_b200.lib import voxel

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'voxel_mesh.npz'))
SHAPES = {'blob32': (32, 32, 32), 'float12': (12, 12, 12), 'full5': (5, 5, 5), 'empty4': (4, 4, 4)}


def _input(name):
    a = G['in_' + name]
    if a.dtype == np.uint8:
        n = int(np.prod(SHAPES[name]))
        return np.unpackbits(a)[:n].astype(bool).reshape(SHAPES[name])
    return a.copy()


@pytest.mark.parametrize('name', sorted(SHAPES))
@pytest.mark.parametrize('surface_view', [True, False])
def test_voxel2obj_matches_reference(name, surface_view, tmp_path):
    key = '%s_sv%d' % (name, surface_view)
    v, f = voxel.voxel2mesh(_input(name), surface_view)
    assert [len(v), len(f)] == G['n_' + key].tolist()
    if 'verts_' + key in G.files:
        assert np.array_equal(v, G['verts_' + key])          # same float64 expression: bit-exact
        assert np.array_equal(f, G['faces_' + key])
    fn = str(tmp_path / 'p.obj')
    voxel.voxel2obj(fn, _input(name), surface_view)
    assert hashlib.sha256(open(fn).read().encode()).hexdigest() == str(G['sha_' + key])


def test_voxel2mesh_marks_input_like_reference():
    vox = np.array([[[0.2, 0.5], [0.31, 0.0]]])
    voxel.voxel2mesh(vox, True)
    assert vox.tolist() == [[[0.2, 1.0], [1.0, 0.0]]]

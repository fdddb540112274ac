"""lib/binvox_rw.py mirror against files written / read by the real reference module
(tests/golden/make_binvox_fixture.py)."""
import io
import os

import numpy as np
import pytest

from r2n2_b200.lib import binvox_rw

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'binvox.npz'))
NAMES = sorted(k[5:] for k in G.files if k.startswith('file_'))


@pytest.mark.parametrize('name', NAMES)
def test_read_matches_reference(name):
    data = G['file_' + name].tobytes()
    meta = G['meta_' + name]
    d = int(meta[0])
    a = binvox_rw.read_as_3d_array(io.BytesIO(data))
    b = binvox_rw.read_as_3d_array(io.BytesIO(data), fix_coords=False)
    assert a.data.dtype == bool and a.data.shape == (d, d, d)
    assert np.array_equal(np.packbits(a.data), G['xyz_' + name])
    assert np.array_equal(np.packbits(b.data), G['xzy_' + name])
    assert a.axis_order == 'xyz' and b.axis_order == 'xzy'
    assert a.dims + a.translate + [a.scale] == meta.tolist()
    c = binvox_rw.read_as_coord_array(io.BytesIO(data))
    assert np.array_equal(binvox_rw.sparse_to_dense(c.data, d), a.data)
    order = np.lexsort(c.data[::-1])                                 # the file lists voxels in x-z-y order
    assert np.array_equal(binvox_rw.dense_to_sparse(a.data), c.data[:, order])


@pytest.mark.parametrize('name', NAMES)
def test_write_is_byte_identical(name):
    data = G['file_' + name].tobytes()
    vox = binvox_rw.read_as_3d_array(io.BytesIO(data))
    out = io.BytesIO()
    vox.write(out)
    assert out.getvalue() == data
    raw = binvox_rw.read_as_3d_array(io.BytesIO(data), fix_coords=False)
    out = io.BytesIO()
    binvox_rw.write(raw, out)
    assert out.getvalue() == data


def test_errors():
    with pytest.raises(IOError):
        binvox_rw.read_as_3d_array(io.BytesIO(b'not a binvox\n'))
    good = G['file_empty8'].tobytes()
    with pytest.raises(IOError):
        binvox_rw.read_as_3d_array(io.BytesIO(good[:-2]))          # runs shorter than the header says
    with pytest.raises(ValueError):
        binvox_rw.Voxels(np.zeros((2, 2, 2), bool), [2, 2, 2], [0, 0, 0], 1.0, 'zyx')

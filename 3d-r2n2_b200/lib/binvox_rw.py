"""Reader / writer for the run-length-encoded .binvox occupancy grids the labels come from
(lib/binvox_rw.py:62-255 of the reference, which is a copy of dimatura/binvox-rw-py).  Host-side file
format code, off the hot path; rewritten for current numpy (the reference module no longer imports:
`np.bool`, `np.int`) and with a writer that emits bytes.

File layout: ASCII header `#binvox 1` / `dim d d d` / `translate x y z` / `scale s` / `data`, then
(value, count) byte pairs, count <= 255, over the grid flattened in x-z-y order."""
import numpy as np


class Voxels(object):
    """data: bool (dims) array in `axis_order` ('xyz' after fix_coords, 'xzy' as stored) or a 3xN coordinate
    array; dims, translate, scale as in the header (lib/binvox_rw.py:62-98)."""

    def __init__(self, data, dims, translate, scale, axis_order):
        if axis_order not in ('xzy', 'xyz'):
            raise ValueError('axis_order must be xzy or xyz')
        self.data = data
        self.dims = dims
        self.translate = translate
        self.scale = scale
        self.axis_order = axis_order

    def clone(self):
        return Voxels(self.data.copy(), self.dims[:], self.translate[:], self.scale, self.axis_order)

    def write(self, fp):
        write(self, fp)


def read_header(fp):
    """lib/binvox_rw.py:101-112; fp is a binary file object."""
    line = fp.readline().strip()
    if not line.startswith(b'#binvox'):
        raise IOError('Not a binvox file')
    dims = [int(i) for i in fp.readline().strip().split(b' ')[1:]]
    translate = [float(i) for i in fp.readline().strip().split(b' ')[1:]]
    scale = [float(i) for i in fp.readline().strip().split(b' ')[1:]][0]
    fp.readline()                                           # 'data'
    return dims, translate, scale


def _runs(fp):
    raw = np.frombuffer(fp.read(), dtype=np.uint8)
    if raw.size % 2:
        raise IOError('binvox: odd number of run-length bytes')
    return raw[::2], raw[1::2]


def read_as_3d_array(fp, fix_coords=True):
    """Dense bool grid (lib/binvox_rw.py:115-137); with fix_coords the stored x-z-y order becomes x-y-z."""
    dims, translate, scale = read_header(fp)
    values, counts = _runs(fp)
    data = np.repeat(values, counts).astype(bool)
    if data.size != int(np.prod(dims)):
        raise IOError('binvox: %d voxels in the runs, header says %s' % (data.size, dims))
    data = data.reshape(dims)
    if fix_coords:
        return Voxels(np.transpose(data, (0, 2, 1)), dims, translate, scale, 'xyz')
    return Voxels(data, dims, translate, scale, 'xzy')


def read_as_coord_array(fp, fix_coords=True):
    """3xN integer coordinates of the occupied voxels (lib/binvox_rw.py:140-185; the reference's true
    divisions are floor divisions here, which is what its Python-2 origin computed)."""
    dims, translate, scale = read_header(fp)
    values, counts = _runs(fp)
    nz = np.flatnonzero(np.repeat(values, counts))
    x = nz // (dims[0] * dims[1])
    zwpy = nz % (dims[0] * dims[1])
    z = zwpy // dims[0]
    y = zwpy % dims[0]
    if fix_coords:
        return Voxels(np.ascontiguousarray(np.vstack((x, y, z))), dims, translate, scale, 'xyz')
    return Voxels(np.ascontiguousarray(np.vstack((x, z, y))), dims, translate, scale, 'xzy')


def dense_to_sparse(voxel_data, dtype=int):
    if voxel_data.ndim != 3:
        raise ValueError('voxel_data is wrong shape; should be 3D array.')
    return np.asarray(np.nonzero(voxel_data), dtype)


def sparse_to_dense(voxel_data, dims, dtype=bool):
    if voxel_data.ndim != 2 or voxel_data.shape[0] != 3:
        raise ValueError('voxel_data is wrong shape; should be 3xN array.')
    if np.isscalar(dims):
        dims = [dims] * 3
    dims = np.atleast_2d(dims).T
    xyz = voxel_data.astype(int)
    xyz = xyz[:, ~np.any((xyz < 0) | (xyz >= dims), 0)]
    out = np.zeros(dims.flatten(), dtype=dtype)
    out[tuple(xyz)] = True
    return out


def write(voxel_model, fp):
    """lib/binvox_rw.py:213-255: same header text and the same greedy run-length pairs (a run is cut at
    255; byte-identical output, including its zero-length pair after a run of exactly k*255), written as bytes to a binary file object."""
    dense = voxel_model.data if voxel_model.data.ndim != 2 else sparse_to_dense(voxel_model.data, voxel_model.dims)
    if voxel_model.axis_order not in ('xzy', 'xyz'):
        raise ValueError('Unsupported voxel model axis order')
    flat = (dense if voxel_model.axis_order == 'xzy' else np.transpose(dense, (0, 2, 1))).flatten().astype(np.uint8)
    fp.write(('#binvox 1\n' + 'dim ' + ' '.join(map(str, voxel_model.dims)) + '\n' +
              'translate ' + ' '.join(map(str, voxel_model.translate)) + '\n' +
              'scale ' + str(voxel_model.scale) + '\ndata\n').encode('ascii'))
    if flat.size == 0:
        return
    starts = np.concatenate(([0], np.flatnonzero(flat[1:] != flat[:-1]) + 1))
    lengths = np.diff(np.concatenate((starts, [flat.size])))
    out = bytearray()
    last = len(starts) - 1
    for i, (v, n) in enumerate(zip(flat[starts].tolist(), lengths.tolist())):
        full, rest = divmod(n, 255)
        out += bytes([v, 255]) * full
        if rest or i != last:        # like the reference, a run of k*255 that ends on a change leaves a (v, 0) pair
            out += bytes([v, rest])
    fp.write(bytes(out))

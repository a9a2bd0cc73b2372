"""Binary STL sink (reference sdf/stl.py:4-24): 80-byte zero header, uint32
triangle count, then 50-byte records {normal f32x3, 3 vertices f32x3, attr u16}."""
import struct

import numpy as np

_RECORD = np.dtype([('normal', '<f4', 3), ('points', '<f4', (3, 3)), ('attr', '<u2')])


def stl_records(points):
    tris = np.asarray(points, dtype='float32').reshape((-1, 3, 3))
    e1 = tris[:, 1] - tris[:, 0]
    e2 = tris[:, 2] - tris[:, 0]
    normals = np.cross(e1, e2)
    with np.errstate(invalid='ignore', divide='ignore'):
        normals /= np.linalg.norm(normals, axis=1).reshape((-1, 1))
    rec = np.zeros(len(tris), dtype=_RECORD)
    rec['points'] = tris
    rec['normal'] = normals
    return rec


def write_binary_stl(path, points):
    rec = stl_records(points)
    with open(path, 'wb') as fp:
        fp.write(b'\x00' * 80)
        fp.write(struct.pack('<I', len(rec)))
        fp.write(rec.tobytes())

"""Indexed-mesh files for ``save()`` when ``meshio`` is not installed (sdf/core.py:152-164).

The reference hands every path that does not end in ``.stl`` to ``meshio.Mesh(points, cells).write(path)``.  With
meshio present this package does the same (core._mesh); without it, the welded mesh coming off the GPU
(``core.weld``: unique vertices + int32 faces, identical to ``np.unique(points, axis=0, return_inverse=True)``) is
written here in the three plain formats people mesh to: Wavefront ``.obj``, Stanford ``.ply`` (binary little
endian, like meshio's default) and ``.off``.  Anything else raises and names meshio.

``IndexedMesh`` mirrors the two attributes and the one method of ``meshio.Mesh`` the reference uses.
"""
import os

import numpy as np


class IndexedMesh:
    def __init__(self, points, cells):
        self.points = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
        self.cells = [(kind, np.ascontiguousarray(conn, dtype=np.int32)) for kind, conn in cells]
        for kind, conn in self.cells:
            if kind != 'triangle' or conn.ndim != 2 or conn.shape[1] != 3:
                raise ValueError('IndexedMesh holds triangle cells (T, 3), got %r %r' % (kind, conn.shape))
            if len(conn) and (conn.min() < 0 or conn.max() >= len(self.points)):
                raise ValueError('face index out of range')

    @property
    def faces(self):
        if not self.cells:
            return np.zeros((0, 3), np.int32)
        return np.concatenate([conn for _, conn in self.cells])

    def write(self, path, file_format=None):
        ext = (file_format or os.path.splitext(path)[1].lstrip('.')).lower()
        writer = _WRITERS.get(ext)
        if writer is None:
            raise ValueError('no built-in writer for %r (built in: %s); install meshio for the other formats'
                             % (ext, ', '.join('.' + e for e in sorted(_WRITERS))))
        writer(path, self.points, self.faces)


def _write_obj(path, verts, faces):
    with open(path, 'w') as fp:
        fp.write('# sdf_b200\n')
        np.savetxt(fp, verts, fmt='v %.9g %.9g %.9g')
        np.savetxt(fp, faces + 1, fmt='f %d %d %d')               # OBJ indices start at 1


def _write_off(path, verts, faces):
    with open(path, 'w') as fp:
        fp.write('OFF\n%d %d 0\n' % (len(verts), len(faces)))
        np.savetxt(fp, verts, fmt='%.9g %.9g %.9g')
        np.savetxt(fp, faces, fmt='3 %d %d %d')


def _write_ply(path, verts, faces):
    header = ('ply\nformat binary_little_endian 1.0\ncomment sdf_b200\n'
              'element vertex %d\nproperty float x\nproperty float y\nproperty float z\n'
              'element face %d\nproperty list uchar int vertex_indices\nend_header\n' % (len(verts), len(faces)))
    rec = np.empty(len(faces), dtype=[('n', 'u1'), ('v', '<i4', (3,))])
    rec['n'] = 3
    rec['v'] = faces
    with open(path, 'wb') as fp:
        fp.write(header.encode('ascii'))
        fp.write(verts.astype('<f4').tobytes())
        fp.write(rec.tobytes())


_WRITERS = {'obj': _write_obj, 'off': _write_off, 'ply': _write_ply}


def read(path):
    """Reads back what ``IndexedMesh.write`` wrote (round-trip checks and small tools, not a general parser)."""
    ext = os.path.splitext(path)[1].lstrip('.').lower()
    if ext == 'obj':
        verts, faces = [], []
        with open(path) as fp:
            for line in fp:
                t = line.split()
                if t and t[0] == 'v':
                    verts.append([float(x) for x in t[1:4]])
                elif t and t[0] == 'f':
                    faces.append([int(x.split('/')[0]) - 1 for x in t[1:4]])
        return IndexedMesh(np.array(verts, np.float32).reshape(-1, 3), [('triangle', np.array(faces, np.int32).reshape(-1, 3))])
    if ext == 'off':
        with open(path) as fp:
            assert fp.readline().strip() == 'OFF'
            nv, nf, _ = (int(x) for x in fp.readline().split())
            data = np.loadtxt(fp, max_rows=nv, ndmin=2) if nv else np.zeros((0, 3))
            conn = np.loadtxt(fp, max_rows=nf, ndmin=2, dtype=np.int64) if nf else np.zeros((0, 4), np.int64)
        return IndexedMesh(data.astype(np.float32), [('triangle', conn[:, 1:4].astype(np.int32))])
    if ext == 'ply':
        with open(path, 'rb') as fp:
            nv = nf = 0
            while True:
                line = fp.readline().decode('ascii').strip()
                if line.startswith('element vertex'):
                    nv = int(line.split()[-1])
                elif line.startswith('element face'):
                    nf = int(line.split()[-1])
                elif line.startswith('format'):
                    assert 'binary_little_endian' in line
                elif line == 'end_header':
                    break
            verts = np.frombuffer(fp.read(12 * nv), '<f4').reshape(-1, 3)
            rec = np.frombuffer(fp.read(13 * nf), dtype=[('n', 'u1'), ('v', '<i4', (3,))])
        return IndexedMesh(verts, [('triangle', rec['v'])])
    raise ValueError('cannot read %r' % path)

"""Vertex welding on the GPU (sdfb200_weld) == np.unique(points, axis=0,
return_inverse=True) -- the reference's core._mesh (sdf/core.py:160-164).  Needs a B200: -m gpu."""
import numpy as np
import pytest

import models

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name,n', [('sphere', 64), ('csg', 128), ('gearlike', 192), ('blobby', 192), ('knurling', 256)])
def test_weld_equals_numpy_unique(name, n):
    import sdf_b200 as S
    from sdf_b200 import core
    f = models.CONFIGS[name](S)
    _, L = models.CONFIG_GRIDS[name]
    pts = S.generate(f, step=2.0 * L / n, bounds=((-L,) * 3, (L,) * 3), verbose=False)
    verts, faces = core.weld(pts)
    ref_v, ref_inv = np.unique(np.asarray(pts), axis=0, return_inverse=True)
    assert verts.shape == ref_v.shape and np.array_equal(verts, ref_v), (name, verts.shape, ref_v.shape)
    assert np.array_equal(faces.reshape(-1), ref_inv.reshape(-1)), name
    assert np.array_equal(verts[faces.reshape(-1)], np.asarray(pts))                 # the indexed mesh is the soup
    # a closed surface: about half as many vertices as triangles
    assert 0.4 < len(verts) / (len(pts) / 3) < 0.6, (len(verts), len(pts) // 3)


def test_weld_edge_cases():
    from sdf_b200 import core
    v, f = core.weld(np.zeros((0, 3), np.float32))
    assert v.shape == (0, 3) and f.shape == (0, 3)
    tri = np.array([[0.0, 0.0, 0.0], [-0.0, 0.0, 1.0], [1.0, -0.0, 0.0],
                    [0.0, -0.0, 1.0], [0.0, 0.0, -0.0], [1.0, 0.0, 0.0]], np.float32)      # -0.0 welds with +0.0
    v, f = core.weld(tri)
    rv, ri = np.unique(tri, axis=0, return_inverse=True)
    assert len(v) == 3 and np.array_equal(v, rv) and np.array_equal(f.reshape(-1), ri.reshape(-1))

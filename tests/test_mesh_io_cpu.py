"""save() to indexed formats without meshio (sdf_b200/mesh_io.py): the files hold exactly the welded mesh.

The welding itself runs on the GPU (tests/test_gpu_weld.py checks it against np.unique); here it is replaced by the
numpy call it stands for (sdf/core.py:160-164), so the writers and core._mesh / core.save can be checked on CPU."""
import os

import numpy as np
import pytest

from sdf_b200 import core, mesh_io


def _soup(n_tri=500, seed=3):
    """A soup with shared vertices: random triangles over a small vertex pool, a few awkward float32 values."""
    rng = np.random.default_rng(seed)
    pool = rng.uniform(-2, 2, (n_tri // 2, 3)).astype(np.float32)
    pool[0] = (0.0, -0.0, 1e-30)
    pool[1] = (np.float32(1) / 3, -1e8, 16777217.0)
    idx = rng.integers(0, len(pool), (n_tri, 3))
    return pool[idx].reshape(-1, 3)


@pytest.fixture()
def host_weld(monkeypatch):
    def weld(points, device=None):
        verts, inverse = np.unique(np.asarray(points, np.float32), axis=0, return_inverse=True)
        return verts, inverse.reshape(-1, 3).astype(np.int32)
    monkeypatch.setattr(core, 'weld', weld)


@pytest.mark.parametrize('ext', ['obj', 'ply', 'off'])
def test_written_file_holds_the_welded_mesh(tmp_path, host_weld, ext):
    pts = _soup()
    mesh = core._mesh(pts)
    if not isinstance(mesh, mesh_io.IndexedMesh):
        pytest.skip('meshio is installed: save() goes through it, as in the reference')
    verts, faces = np.unique(pts, axis=0, return_inverse=True)
    assert np.array_equal(mesh.points, verts) and np.array_equal(mesh.cells[0][1], faces.reshape(-1, 3))
    path = os.path.join(str(tmp_path), 'm.' + ext)
    mesh.write(path)
    back = mesh_io.read(path)
    assert back.points.dtype == np.float32 and back.faces.dtype == np.int32
    assert np.array_equal(back.points, verts)                            # every float32 survives '%.9g'
    assert np.array_equal(back.faces, faces.reshape(-1, 3))
    assert np.array_equal(back.points[back.faces].reshape(-1, 3), pts)   # the soup again


def test_save_dispatches_on_the_extension(tmp_path, host_weld, monkeypatch):
    pts = _soup(60)
    monkeypatch.setattr(core, 'generate', lambda *a, **k: pts)
    if not isinstance(core._mesh(pts), mesh_io.IndexedMesh):
        pytest.skip('meshio is installed')
    path = os.path.join(str(tmp_path), 'out.PLY')
    core.save(path, object())
    assert len(mesh_io.read(path).faces) == 60
    with pytest.raises(ValueError, match='install meshio'):
        core.save(os.path.join(str(tmp_path), 'out.vtk'), object())


def test_empty_mesh_and_bad_cells(tmp_path):
    m = mesh_io.IndexedMesh(np.zeros((0, 3), np.float32), [('triangle', np.zeros((0, 3), np.int32))])
    for ext in ('obj', 'ply', 'off'):
        path = os.path.join(str(tmp_path), 'e.' + ext)
        m.write(path)
        back = mesh_io.read(path)
        assert len(back.points) == 0 and len(back.faces) == 0
    with pytest.raises(ValueError):
        mesh_io.IndexedMesh(np.zeros((3, 3)), [('quad', np.zeros((1, 4), np.int32))])
    with pytest.raises(ValueError):
        mesh_io.IndexedMesh(np.zeros((3, 3)), [('triangle', np.array([[0, 1, 3]]))])

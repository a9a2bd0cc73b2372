"""CPU suite, part 1: the oracle is pinned against vectors produced by the REAL reference
(tests/golden/make_golden.py imports /root/reference/sdf with a stub skimage)."""
import numpy as np
import pytest

import models
import sdf_b200
from oracle import sdf_ref
from util_cmp import mesh_topology


@pytest.fixture(scope='module')
def corpus():
    return models.build(sdf_b200)


def test_node_evaluation_bit_identical_to_reference(corpus, golden_dir):
    g = np.load(golden_dir + '/node_eval.npz')
    P = g['points']
    assert len(corpus) >= 50
    for name, f in corpus.items():
        with np.errstate(all='ignore'):
            d = sdf_ref.evaluate(f, P).reshape(-1)
        assert np.array_equal(d, g['d_' + name], equal_nan=True), name


def test_random_compositions_bit_identical_to_reference(golden_dir):
    """120 seeded random trees (depth <= 4: transforms, warps, arrays, smooth folds, 2-D -> 3-D) evaluated
    by the REAL reference (tests/golden/random_trees.npz): the restatement reproduces every value."""
    g = np.load(golden_dir + '/random_trees.npz')
    P = g['points']
    trees = models.random_trees(sdf_b200)
    assert len(trees) == 120
    for i, f in enumerate(trees):
        with np.errstate(all='ignore'):
            d = sdf_ref.evaluate(f, P).reshape(-1)
        assert np.array_equal(d, g['d_%03d' % i], equal_nan=True), i


def test_natural_grid_classification_matches_reference(golden_dir):
    """SURVEY.md section 8(c) table row 3 at FULL size (gearlike, samples = 512^3, the reference's own bounds:
    732 x 732 x 251 samples, 4,232 batches): the oracle's skipped / empty / nonempty equal the real
    reference's (tests/golden/natural.npz).  ~20 s on 8 threads."""
    e, g = np.load(golden_dir + '/engine.npz'), np.load(golden_dir + '/natural.npz')
    f = models.CONFIGS['gearlike'](sdf_b200)
    lo, hi = e['bounds_gearlike']
    pts, st = sdf_ref.generate(f, step=float(g['autostep_gearlike'][0]), bounds=(tuple(lo), tuple(hi)), workers=8)
    gold = g['auto_gearlike']
    assert st['dims'] == tuple(gold[4:7])
    assert (st['skipped'], st['empty'], st['nonempty']) == tuple(gold[:3])
    assert st['triangles'] == models.NATURAL_TRIANGLES['gearlike'] and len(pts) == 3 * st['triangles']


def test_known_answers():
    # SURVEY.md section 8c golden vectors
    s = sdf_b200.sphere(1)
    assert sdf_ref.evaluate(s, np.array([(0., 0, 0), (1, 0, 0), (2, 0, 0)])).reshape(-1).tolist() == [-1.0, 0.0, 1.0]
    c = sdf_b200.cylinder(0.5).orient(sdf_b200.X)
    d = sdf_ref.evaluate(c, np.array([(0, .6, 0), (1, 0, .6)])).reshape(-1)
    assert np.allclose(d, [0.1, 0.1], atol=1e-15)


def test_estimate_bounds_identical_to_reference(golden_dir):
    g = np.load(golden_dir + '/engine.npz')
    for name in models.CONFIGS:
        f = models.CONFIGS[name](sdf_b200)
        b = sdf_ref.estimate_bounds(f)
        assert np.array_equal(np.array(b), g['bounds_' + name]), name
    # the value quoted in SURVEY.md section 3.4
    assert g['bounds_sphere'][0][0] == -1.0567876661951774 and g['bounds_sphere'][1][0] == 1.0567883957482447


@pytest.mark.parametrize('name,n', [('sphere', 64), ('csg', 96), ('gearlike', 128), ('blobby', 128)])
def test_batch_classification_matches_reference(golden_dir, name, n):
    g = np.load(golden_dir + '/engine.npz')
    f = models.CONFIGS[name](sdf_b200)
    _, L = models.CONFIG_GRIDS[name]
    pts, st = sdf_ref.generate(f, step=2.0 * L / n, bounds=((-L, -L, -L), (L, L, L)))
    gold = g['cls_%s_%d' % (name, n)]
    assert (st['skipped'], st['empty'], st['nonempty']) == tuple(gold[:3])
    assert st['dims'] == tuple(gold[4:7])
    # self-consistency that stands in for the missing skimage oracle: closed, oriented 2-manifold
    topo = mesh_topology(pts)
    assert topo['boundary'] == 0 and topo['nonmanifold'] == 0 and topo['repeated_directed'] == 0
    expected_chi = {'sphere': 2, 'csg': -8, 'gearlike': 0, 'blobby': 2}[name]
    assert topo['V'] - topo['E'] + topo['F'] == expected_chi


def test_sphere_area_and_volume_converge():
    f = sdf_b200.sphere(1)
    errs = []
    for n in (32, 64):
        pts, _ = sdf_ref.generate(f, step=2.5 / n, bounds=((-1.25,) * 3, (1.25,) * 3))
        t = pts.reshape(-1, 3, 3)
        area = 0.5 * np.linalg.norm(np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]), axis=1).sum()
        vol = np.einsum('ij,ij->i', t[:, 0], np.cross(t[:, 1], t[:, 2])).sum() / 6   # outward normals -> positive
        errs.append((abs(area - 4 * np.pi), abs(vol - 4 * np.pi / 3)))
        assert vol > 0
    assert errs[1][0] < errs[0][0] / 3 and errs[1][1] < errs[0][1] / 3      # O(step^2)
    assert errs[1][0] < 1e-2 and errs[1][1] < 1e-2


def test_threaded_generate_equals_serial():
    f = models.cfg_csg(sdf_b200)
    a, sa = sdf_ref.generate(f, step=2.0 / 64, bounds=((-1,) * 3, (1,) * 3))
    b, sb = sdf_ref.generate(f, step=2.0 / 64, bounds=((-1,) * 3, (1,) * 3), workers=4)
    assert np.array_equal(a, b) and sa == sb


def test_marching_cubes_error_paths():
    # skimage raises ValueError / RuntimeError; the reference maps both to an empty batch (core.py:53-56)
    with pytest.raises(sdf_ref.MCValueError):
        sdf_ref.marching_cubes_soup(np.ones((4, 4, 4)))                # level outside the data range
    with pytest.raises(sdf_ref.MCValueError):
        sdf_ref.marching_cubes_soup(np.zeros((1, 4, 4)))               # a dimension < 2
    v = -np.ones((3, 3, 3)); v[1, 1, 1] = 0.0                          # max == level: in range, but no v > level
    with pytest.raises(sdf_ref.MCRuntimeError):
        sdf_ref.marching_cubes_soup(v)
    v[1, 1, 1] = 1.0
    soup = sdf_ref.marching_cubes_soup(v)
    assert len(soup) == 3 * 8 and soup.dtype == np.float32              # an octahedron around the centre sample

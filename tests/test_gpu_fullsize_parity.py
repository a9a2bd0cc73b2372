"""Vertex-level parity of the CUDA path against the CPU oracle AT THE SIZES BASELINE.json STATES
(SURVEY.md section 8d "primary" lattices): sphere 64^3 as one dense batch, the CSG example at 256^3,
gearlike at 512^3, blobby at 768^3 and -- the oracle's numpy tree being slow on 1.07 G points -- three
seeded Z batch layers (3 x 1024 = 3072 batches) of knurling at 1024^3.

Bar (north_star): identical triangle count, vertices within 1e-5 after sort; here additionally >= 99.9 %
of the triangles bit-identical.  What the reference does per batch: core.py:45-60.  Needs a B200: -m gpu.
"""
import os

import numpy as np
import pytest

import models
from util_cmp import compare_soups

pytestmark = [pytest.mark.gpu, pytest.mark.fullsize]

BATCH = 32


def _check(got_pts, got_info, ref_pts, ref_stats, label):
    assert (got_info['skipped'], got_info['empty'], got_info['nonempty']) == (
        ref_stats['skipped'], ref_stats['empty'], ref_stats['nonempty']), (label, got_info, ref_stats)
    rep = compare_soups(got_pts, ref_pts)
    assert rep['count_equal'], (label, rep)
    assert rep['max_vertex_dist'] <= 1e-5, (label, rep)
    assert rep['exact_triangles'] >= 0.999 * rep['n_ref'], (label, rep)
    return rep


def test_config1_sphere_64_single_dense_batch():
    """configs[0]: sphere(1) on a 64^3 lattice, sparse=False, batch_size=64 -> literally one batch (core.py:114-122)."""
    import sdf_b200 as sdf
    from oracle import sdf_ref
    f = models.cfg_sphere(sdf)
    n, L = models.CONFIG_GRIDS['sphere']
    kw = dict(step=2.0 * L / n, bounds=((-L, -L, -L), (L, L, L)), batch_size=64, sparse=False)
    pts, info = sdf.generate(f, verbose=False, return_info=True, workers=1, **kw)
    ref_pts, ref_stats = sdf_ref.generate(f, **kw)
    assert info['batches'] == 1 and ref_stats['batches'] == 1
    _check(pts, info, ref_pts, ref_stats, 'sphere64')


@pytest.mark.parametrize('name', ['csg', 'gearlike', 'blobby'])
def test_full_lattice_matches_oracle(name):
    """configs[1..3] at their stated sizes; the oracle runs the reference's ThreadPool over all host threads."""
    import sdf_b200 as sdf
    from oracle import sdf_ref
    f = models.CONFIGS[name](sdf)
    n, L = models.CONFIG_GRIDS[name]
    kw = dict(step=2.0 * L / n, bounds=((-L, -L, -L), (L, L, L)))
    pts, info = sdf.generate(f, verbose=False, return_info=True, **kw)
    assert info['dims'] == (n, n, n)
    ref_pts, ref_stats = sdf_ref.generate(f, workers=os.cpu_count() or 1, **kw)
    _check(pts, info, ref_pts, ref_stats, name)


def test_knurling_1024_seeded_layers_match_oracle():
    """configs[4]: three seeded Z batch layers of the 1024^3 lattice (3072 of 32768 batches); the GPU runs each
    layer through the Z-slab range of the C-ABI, the oracle the same batches (core.py:119 batch order)."""
    import sdf_b200 as sdf
    from sdf_b200 import core
    from oracle import sdf_ref
    f = models.cfg_knurling(sdf)
    n, L = models.CONFIG_GRIDS['knurling']
    step, bounds = 2.0 * L / n, ((-L, -L, -L), (L, L, L))
    _, _, (X, Y, Z) = core.make_lattice(f, step, bounds)
    assert (len(X), len(Y), len(Z)) == (n, n, n)
    nb = n // BATCH
    layers = np.sort(np.random.default_rng(1024).choice(np.arange(2, nb - 2), 3, replace=False))
    for bz in layers:
        tris, info = core.run_batches(f, X, Y, Z, BATCH, True, 'exact', zb=(int(bz), int(bz) + 1))
        subset = [(bx * nb + by) * nb + int(bz) for bx in range(nb) for by in range(nb)]
        ref_pts, ref_stats = sdf_ref.generate(f, step=step, bounds=bounds, batch_subset=subset,
                                              workers=os.cpu_count() or 1)
        rep = _check(np.asarray(tris).reshape(-1, 3), info, ref_pts, ref_stats, 'knurling layer %d' % bz)
        assert rep['n_ref'] > 0

"""SDF3.__call__ on the GPU (float64) for 120 seeded random compositions against values produced by the
REAL reference (tests/golden/random_trees.npz).  Needs a B200: -m gpu.  (Named to run last.)

NaN semantics: numpy's minimum / maximum (dn.py:14,29,44) propagate a NaN operand (the only NaN source in these
trees is the exact centre of an `ellipsoid`, 0/0, d3.py:255-258).  Round 1's device code used fmin/fmax, which
drop it (trees 43 and 116 at the origin); the accurate math policies now follow numpy (sdf_ops.cuh SDF_MINMAX)
and the NaN positions must agree exactly.
"""
import numpy as np
import pytest

import models

pytestmark = pytest.mark.gpu

KNOWN_NAN_DROPPED = set()          # (tree, point) with a reference NaN and a finite device value: none allowed


@pytest.mark.parametrize('front', ['interp', 'jit'])
def test_eval_f64_random_trees_match_reference_goldens(front, golden_dir):
    import os
    import sdf_b200 as sdf
    old = os.environ.get('SDFB200_JIT')
    os.environ['SDFB200_JIT'] = '0' if front == 'interp' else '2'
    try:
        g = np.load(golden_dir + '/random_trees.npz')
        P = g['points']
        trees = models.random_trees(sdf)
        step = 1 if front == 'interp' else 6         # one NVRTC compile per tree on the specialised path
        dropped = set()
        for i in range(0, len(trees), step):
            d = trees[i](P).reshape(-1)
            r = g['d_%03d' % i]
            assert not np.any(np.isnan(d) & ~np.isnan(r)), i          # never NaN where the reference is finite
            dropped |= {(i, int(k)) for k in np.nonzero(np.isnan(r) & ~np.isnan(d))[0]}
            ok = np.isfinite(r) & np.isfinite(d)
            err = np.max(np.abs(d[ok] - r[ok]) / np.maximum(1.0, np.abs(r[ok]))) if ok.any() else 0.0
            assert err <= 1e-11, (i, err)
        assert dropped <= KNOWN_NAN_DROPPED, dropped
    finally:
        if old is None:
            os.environ.pop('SDFB200_JIT', None)
        else:
            os.environ['SDFB200_JIT'] = old


@pytest.mark.parametrize('name,n', [('gearlike', 512), ('blobby', 768), ('knurling', 1024)])
def test_natural_grids_match_the_reference_classification(name, n, golden_dir):
    """SURVEY.md section 8(c) table, rows 3-5, at FULL size: `samples = N^3` with the bounds estimate on the
    GPU evaluator; lattice dimensions, skipped / empty / nonempty batches and the number of active cubes
    equal what the REAL reference computes (tests/golden/natural.npz, `make_golden.py --natural`)."""
    import sdf_b200 as sdf
    g = np.load(golden_dir + '/natural.npz')
    f = models.CONFIGS[name](sdf)
    pts, info = sdf.generate(f, samples=n ** 3, verbose=False, return_info=True)
    gold = g['auto_' + name]
    assert info['dims'] == tuple(gold[4:7]), (info['dims'], gold)
    assert (info['skipped'], info['empty'], info['nonempty']) == tuple(gold[:3]), (info, gold)
    assert info['active_cubes'] == gold[3], (info['active_cubes'], gold[3])
    assert len(pts) == 3 * info['triangles'] and info['triangles'] >= info['active_cubes']
    if name in models.NATURAL_TRIANGLES:                      # identical triangle count (north_star's bar) with the oracle
        assert info['triangles'] == models.NATURAL_TRIANGLES[name], (info['triangles'], models.NATURAL_TRIANGLES[name])

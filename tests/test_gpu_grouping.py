"""The two speed-only mechanisms of the tile pipeline must never change the mesh (needs a B200: -m gpu):

* sample grouping of the specialised tile kernel (axis sharing, sdfb200_jit_plan): whatever axis and
  group size, 'exact' output is bit-identical;
* pipelined host output (SDFB200_PIPE): triangles leave in rounds while the next round computes.
"""
import os

import numpy as np
import pytest

import models
from util_cmp import canonical_triangles

pytestmark = pytest.mark.gpu


@pytest.fixture()
def env():
    keys = ('SDFB200_JIT', 'SDFB200_JIT_SHARE', 'SDFB200_JIT_N', 'SDFB200_PIPE')
    old = {k: os.environ.get(k) for k in keys}
    yield os.environ
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _fresh_generate(sdf, f, **kw):
    from sdf_b200 import core
    core._handles.clear()                 # the grouping is chosen when a handle compiles its kernels
    return sdf.generate(f, verbose=False, return_info=True, **kw)


def _models(sdf):
    return {
        'gearlike': (models.CONFIGS['gearlike'](sdf), 2.1),
        'cyl_x': (sdf.cylinder(0.7).orient(sdf.X) & sdf.sphere(1.5), 1.6),
        'cyl_y': (sdf.rounded_box((2, 1, 1.5), 0.2) - sdf.cylinder(0.4).orient(sdf.Y).k(0.1), 1.3),
        'knurling': (models.CONFIGS['knurling'](sdf), 3.0),
    }


@pytest.mark.parametrize('name', ['gearlike', 'cyl_x', 'cyl_y', 'knurling'])
def test_sample_grouping_does_not_change_the_mesh(env, name):
    import sdf_b200 as sdf
    f, L = _models(sdf)[name]
    n = 150                                # ragged: 150 = 4 * 32 + 22 -> partial tiles on every axis
    bounds, step = ((-L, -L, -L * 0.9), (L, L * 0.8, L)), 2.0 * L / n
    env['SDFB200_JIT'] = '2'
    env['SDFB200_JIT_SHARE'] = '-1'; env['SDFB200_JIT_N'] = '4'
    base, ib = _fresh_generate(sdf, f, step=step, bounds=bounds, precision='exact')
    base = canonical_triangles(base)
    assert len(base) > 1000
    for share, group in ((0, 11), (1, 6), (2, 11), (2, 3), (1, 16), (0, 1)):
        env['SDFB200_JIT_SHARE'] = str(share); env['SDFB200_JIT_N'] = str(group)
        got, ig = _fresh_generate(sdf, f, step=step, bounds=bounds, precision='exact')
        assert (ig['skipped'], ig['empty'], ig['nonempty'], ig['active_cubes']) == \
               (ib['skipped'], ib['empty'], ib['nonempty'], ib['active_cubes']), (share, group)
        assert ig['evaluated'] == ib['evaluated']
        assert np.array_equal(canonical_triangles(got), base), (share, group)
    # the automatic choice (what users get)
    env.pop('SDFB200_JIT_SHARE'); env.pop('SDFB200_JIT_N')
    got, _ = _fresh_generate(sdf, f, step=step, bounds=bounds, precision='exact')
    assert np.array_equal(canonical_triangles(got), base)
    # 'f32' keeps float32 values: the grouping may move them by an ulp (FMA contraction), not more
    env['SDFB200_JIT_SHARE'] = '-1'; env['SDFB200_JIT_N'] = '4'
    a, _ = _fresh_generate(sdf, f, step=step, bounds=bounds, precision='f32')
    env['SDFB200_JIT_SHARE'] = '2'; env['SDFB200_JIT_N'] = '11'
    b, _ = _fresh_generate(sdf, f, step=step, bounds=bounds, precision='f32')
    assert abs(len(a) - len(b)) <= 0.001 * len(a)


@pytest.mark.parametrize('front,n,sparse', [('0', 224, False), ('2', 224, False), ('2', 416, True)])
def test_pipelined_host_output_equals_single_copy(env, front, n, sparse):
    """Rounds are whole waves of 148 tiles and a job needs two waves to be cut at all: 224^3 with the
    skip test off = 343 tiles (3 rounds), 416^3 sparse = a few hundred active tiles."""
    import sdf_b200 as sdf
    f = models.CONFIGS['gearlike'](sdf)
    L = 2.1
    bounds, step = ((-L, -L, -L), (L, L, L)), 2.0 * L / n
    env['SDFB200_JIT'] = front
    env['SDFB200_PIPE'] = '1'
    base, ib = _fresh_generate(sdf, f, step=step, bounds=bounds, precision='exact', sparse=sparse)
    assert ib['evaluated'] >= 2 * 148 * 33 ** 3 * 0.9, 'lattice too small to exercise the pipelined path'
    base = canonical_triangles(base)
    assert len(base) > 100000
    for rounds in ('2', '3', '16'):
        env['SDFB200_PIPE'] = rounds
        got, ig = _fresh_generate(sdf, f, step=step, bounds=bounds, precision='exact', sparse=sparse)
        assert (ig['skipped'], ig['empty'], ig['nonempty'], ig['active_cubes'], ig['evaluated']) == \
               (ib['skipped'], ib['empty'], ib['nonempty'], ib['active_cubes'], ib['evaluated']), rounds
        assert np.array_equal(canonical_triangles(got), base), rounds
    # a buffer that is too small comes back as "capacity" (nothing truncated), core.run_batches retries
    env['SDFB200_PIPE'] = '4'
    from sdf_b200 import core
    old = core.default_capacity
    core.default_capacity = lambda X, Y, Z: 1000
    core._capacity_hint.clear()
    try:
        got, _ = _fresh_generate(sdf, f, step=step, bounds=bounds, precision='exact', sparse=sparse)
    finally:
        core.default_capacity = old
        core._capacity_hint.clear()
    assert np.array_equal(canonical_triangles(got), base)

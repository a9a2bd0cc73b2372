"""Precision `exact` (fast float32 classification + float64 for whatever can reach the output) must equal precision
`f64` (the reference arithmetic for every sample) ALSO on trees that amplify float32 rounding errors: the guard
band scales with the tree's a-priori gain (sdf_b200/errbound.py; ADVICE round 1: "tau is validated only on the fixed
model corpus").  Needs a B200: -m gpu."""
import numpy as np
import pytest

from util_cmp import compare_soups

pytestmark = pytest.mark.gpu


def _trees(S):
    return {
        'bend_radial_thin_band': S.box((1.2, 1.2, 0.6)).bend_radial(0.4, 0.401, 0.35),
        'transition_radial_thin_band': S.box(1).transition_radial(S.sphere(0.7), 0.5, 0.5005),
        'twist_fast': S.box((0.8, 0.8, 1.6)).twist(40.0),
        'elastic_bend': S.box((1.4, 0.5, 0.5)).bend_linear(S.X * -0.7, S.X * 0.7, S.Z * 0.3, S.ease.in_out_elastic),
        'square_jump': S.box((1.4, 0.5, 0.5)).bend_linear(S.X * -0.7, S.X * 0.7, S.Z * 0.3, S.ease.in_out_square),
        'scaled_up': S.sphere(1e-3).scale(800.0) - S.box(0.9),
    }


@pytest.mark.parametrize('name', list(_trees(__import__('sdf_b200'))))
def test_exact_equals_f64_on_amplifying_trees(name):
    import sdf_b200 as S
    f = _trees(S)[name]
    kw = dict(step=2.4 / 96, bounds=((-1.2, -1.2, -1.2), (1.2, 1.2, 1.2)), verbose=False, return_info=True)
    a, ia = S.generate(f, precision='exact', **kw)
    b, ib = S.generate(f, precision='f64', **kw)
    assert (ia['skipped'], ia['empty'], ia['nonempty']) == (ib['skipped'], ib['empty'], ib['nonempty']), (name, ia, ib)
    rep = compare_soups(a, b)
    assert rep['count_equal'] and rep['exact_triangles'] == rep['n_ref'], (name, rep)
    assert ia['triangles'] > 0

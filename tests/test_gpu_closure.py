"""A tree that holds a user-written numpy closure (reference README "writing your own SDFs") is meshed through the
batch-by-batch route: the closure evaluated on the host as the reference does (library sub-trees still on the GPU),
marching cubes + world mapping on the device (sdfb200_mc_volume).  ADVICE round 1: such trees used to raise.
Needs a B200: -m gpu."""
import os

import numpy as np
import pytest

from util_cmp import compare_soups

pytestmark = pytest.mark.gpu


def _pair():
    import sdf_b200 as S

    @S.sdf3
    def my_sphere(r):
        def f(p):
            return np.linalg.norm(p, axis=1) - r
        return f
    cut = S.box(0.9).translate((0.45, 0.3, 0.2)).rotate(0.4, S.Z)
    return my_sphere(0.8) - cut.k(0.05), S.sphere(0.8) - cut.k(0.05)


def test_generate_with_a_python_closure_matches_the_library_tree(tmp_path):
    import sdf_b200 as S
    mine, lib = _pair()
    kw = dict(step=2.0 / 80, bounds=((-1, -1, -1), (1, 1, 1)), verbose=False, return_info=True)
    a, ia = S.generate(mine, **kw)
    b, ib = S.generate(lib, precision='f64', **kw)
    assert (ia['skipped'], ia['empty'], ia['nonempty']) == (ib['skipped'], ib['empty'], ib['nonempty']), (ia, ib)
    rep = compare_soups(a, b)
    assert rep['count_equal'] and rep['max_vertex_dist'] <= 1e-5 and rep['exact_triangles'] >= 0.999 * rep['n_ref'], rep
    path = os.path.join(str(tmp_path), 'closure.stl')
    mine.save(path, step=2.0 / 40, bounds=((-1, -1, -1), (1, 1, 1)), verbose=False)
    assert os.path.getsize(path) > 84 and (os.path.getsize(path) - 84) % 50 == 0


def test_mc_volume_equals_the_pipeline_on_a_library_batch():
    """sdfb200_mc_volume on a batch volume == generate() on the same single-batch lattice."""
    import sdf_b200 as S
    from sdf_b200 import core
    f = S.sphere(0.7) | S.box(0.5).translate((0.4, 0.2, 0.1))
    X = np.arange(-1.0, 1.0, 2.0 / 33)[:33]
    P = core._cartesian_product(X, X, X)
    vol = f(P).reshape(33, 33, 33)
    h, _ = core.program_handle(f)
    tris = h.mc_volume(vol, X, X, X)
    ref, info = core.run_batches(f, X, X, X, 32, False, 'f64')
    rep = compare_soups(np.asarray(tris).reshape(-1, 3), np.asarray(ref).reshape(-1, 3))
    assert rep['count_equal'] and rep['exact_triangles'] == rep['n_ref'] and rep['n_ref'] > 0, rep


def test_mc_volume_on_nan_infinite_and_denormal_samples_equals_the_oracle():
    """Caller-supplied volumes (a user's closure can return anything): NaN, +-inf, +-0, huge and tiny values.  The extraction
    kernel gives the C restatement's triangles bit for bit -- a vertex on an edge with a NaN corner (or two infinite
    ones) is NaN in all three coordinates, as skimage's generic weighted form makes it.  (The difference -- only the
    coordinate along the edge was NaN -- was found by executing the kernel on the host, tests/hostsim/.)"""
    import sdf_b200 as S
    from oracle import sdf_ref
    from sdf_b200 import core
    from util_cmp import soup_multiset_bits, special_value_volumes
    h, _ = core.program_handle(S.sphere(1))
    ax = np.arange(12.0)
    for vol in special_value_volumes(denormals=False):
        got = h.mc_volume(vol, ax, ax, ax)
        want = sdf_ref.marching_cubes_soup(vol, 0.0)
        assert soup_multiset_bits(got) == soup_multiset_bits(want) and len(got) > 3000

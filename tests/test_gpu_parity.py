"""Parity of the CUDA path (through the C-ABI) against the CPU oracle.  Needs a B200: -m gpu."""
import numpy as np
import pytest

import models
from util_cmp import compare_soups, mesh_topology

pytestmark = pytest.mark.gpu

SMALL = [('sphere', 64), ('csg', 96), ('gearlike', 128), ('blobby', 128), ('knurling', 160)]


@pytest.fixture(scope='module', params=['interp', 'jit'], autouse=True)
def evaluator_front_end(request):
    """Every test runs twice: on the interpreter kernels (SDFB200_JIT=0) and on the run-time
    specialised kernels (SDFB200_JIT=2: always compile, whatever the job size)."""
    import os
    old = os.environ.get('SDFB200_JIT')
    os.environ['SDFB200_JIT'] = '0' if request.param == 'interp' else '2'
    yield request.param
    if old is None:
        os.environ.pop('SDFB200_JIT', None)
    else:
        os.environ['SDFB200_JIT'] = old


@pytest.fixture(scope='module')
def sdf():
    import sdf_b200
    from sdf_b200 import _abi
    assert _abi.device_count() >= 1, 'no CUDA device: the product path has no CPU fallback'
    return sdf_b200


@pytest.fixture(scope='module')
def ref():
    from oracle import sdf_ref
    return sdf_ref


JIT_MUST = ('cfg_', 'weave', 'pawn', 'slice', 'extrude_to', 'wrap_around', 'x_polygon', 'bend_', 'repeat', 'circular')


@pytest.fixture(scope='module')
def corpus(sdf, evaluator_front_end):
    """All 114 models on the interpreter; every third one plus the structurally interesting ones on
    the run-time specialised kernels (each costs three NVRTC compiles)."""
    allm = models.build(sdf)
    if evaluator_front_end == 'interp':
        return allm
    keep = {k: v for i, (k, v) in enumerate(allm.items()) if i % 3 == 0 or any(m in k for m in JIT_MUST)}
    return keep


def test_eval_f64_matches_reference_goldens(sdf, corpus, golden_dir):
    """SDF3.__call__ on the GPU (float64) against vectors produced by the REAL reference."""
    g = np.load(golden_dir + '/node_eval.npz')
    P = g['points']
    worst = {}
    for name, f in corpus.items():
        d = f(P).reshape(-1)
        r = g['d_' + name]
        ok = np.isfinite(r)
        scale = np.maximum(1.0, np.abs(r[ok]))
        err = np.max(np.abs(d[ok] - r[ok]) / scale) if ok.any() else 0.0
        worst[name] = err
        assert np.array_equal(np.isnan(d), np.isnan(r)), name
    bad = {k: v for k, v in worst.items() if v > 5e-13}
    assert not bad, bad


def test_eval_f32_stream_matches_oracle(sdf, ref, corpus):
    """The streamed float32 evaluator (TMA-staged float3 points) against the float64 oracle."""
    import torch
    P64 = models.random_points(1 << 16, seed=99)
    P32 = P64.astype(np.float32)
    Pd = torch.from_numpy(P32).cuda()
    for name, f in corpus.items():
        d = f(Pd).cpu().numpy().reshape(-1)
        r = ref.evaluate(f, P32.astype(np.float64)).reshape(-1)
        ok = np.isfinite(r) & np.isfinite(d)
        assert ok.mean() > 0.99, name
        # tolerance: 256 float32 ulps of max(1, |p|, |d|); the trees amplify input rounding
        tol = 256 * np.finfo(np.float32).eps * np.maximum(1.0, np.maximum(np.abs(r[ok]), 3.0))
        err = np.abs(d[ok] - r[ok])
        assert np.all(err <= tol), (name, float(err.max()), float(tol.min()))


def test_eval_ragged_sizes(sdf, ref):
    """Empty, tiny and non-multiple-of-tile point blocks; host and device buffers."""
    import torch
    f = models.cfg_csg(sdf)
    for n in (0, 1, 3, 1023, 1024, 1025, 4099):
        P = models.random_points(max(n, 8), seed=n)[:n]
        d = f(P)
        assert d.shape == (n, 1)
        if n:
            r = ref.evaluate(f, P)
            assert np.max(np.abs(d - r)) < 1e-12
            dd = f(torch.from_numpy(P.astype(np.float32)).cuda()).cpu().numpy()
            assert np.max(np.abs(dd.reshape(-1) - r.reshape(-1))) < 1e-4


@pytest.mark.parametrize('name,n', SMALL)
@pytest.mark.parametrize('precision', ['f64', 'mixed', 'exact'])
def test_generate_matches_oracle(sdf, ref, golden_dir, name, n, precision):
    """Whole hot path: batch classification == reference goldens, triangles == oracle."""
    f = models.CONFIGS[name](sdf)
    _, L = models.CONFIG_GRIDS[name]
    bounds, step = ((-L, -L, -L), (L, L, L)), 2.0 * L / n
    pts, info = sdf.generate(f, step=step, bounds=bounds, verbose=False, precision=precision, return_info=True)
    gold = np.load(golden_dir + '/engine.npz')['cls_%s_%d' % (name, n)]
    assert (info['skipped'], info['empty'], info['nonempty']) == tuple(gold[:3])
    assert info['active_cubes'] == gold[3]
    assert info['dims'] == tuple(gold[4:7])
    rpts, rstats = ref.generate(f, step=step, bounds=bounds)
    rep = compare_soups(pts, rpts)
    assert rep['count_equal'], rep                       # identical triangle count
    assert rep['max_vertex_dist'] <= 1e-5, rep           # north_star tolerance
    if precision == 'f64':
        # same float64 arithmetic up to libm/FMA differences: expect (almost) every triangle bit-equal
        assert rep['exact_triangles'] >= 0.999 * rep['n_ref'], rep
    if precision != 'mixed':
        # bit-exact watertightness across tile boundaries needs bit-identical values on the shared
        # planes: guaranteed when every output-relevant sample is float64-derived ('exact', 'f64');
        # 'mixed' keeps float32 values, which may differ by an ulp between the two tiles' code paths
        topo = mesh_topology(pts)
        assert topo['boundary'] == 0 and topo['nonmanifold'] == 0 and topo['repeated_directed'] == 0, topo


@pytest.mark.parametrize('name,n', SMALL)
def test_exact_mode_equals_f64_mode_bitwise(sdf, name, n):
    """PREC_EXACT (fast float32 classification + float64 on everything that reaches the output)
    must reproduce the all-float64 run triangle for triangle."""
    f = models.CONFIGS[name](sdf)
    _, L = models.CONFIG_GRIDS[name]
    bounds, step = ((-L, -L, -L), (L, L, L)), 2.0 * L / n
    a, ia = sdf.generate(f, step=step, bounds=bounds, verbose=False, precision='f64', return_info=True)
    b, ib = sdf.generate(f, step=step, bounds=bounds, verbose=False, precision='exact', return_info=True)
    assert (ia['skipped'], ia['empty'], ia['nonempty'], ia['active_cubes']) == (ib['skipped'], ib['empty'], ib['nonempty'], ib['active_cubes'])
    rep = compare_soups(b, a)
    assert rep['count_equal'] and rep['exact_triangles'] == rep['n_ref'], rep
    assert ib['refined'] < 0.2 * ib['evaluated']


def test_fast_stream_evaluator_error_is_inside_the_guard_band(sdf, ref, corpus):
    """The approximate-math evaluator may only be wrong by less than tau (the samples it gets
    'wrong enough to matter' are exactly the ones PREC_EXACT re-evaluates in float64)."""
    import torch
    from sdf_b200 import core
    P64 = models.random_points(1 << 16, seed=5, L=2.5)
    P32 = P64.astype(np.float32)
    Pd = torch.from_numpy(P32).cuda()
    X = np.array([-2.5, 2.5])
    for name, f in corpus.items():
        h, _ = core.program_handle(f)
        out = torch.empty(len(P32), device='cuda')
        h.eval_f32_fast(Pd, out, len(P32))
        d = out.cpu().numpy()
        r = ref.evaluate(f, P32.astype(np.float64)).reshape(-1)
        ok = np.isfinite(r) & np.isfinite(d) & (np.abs(r) < 1.0)
        tau = core.mixed_tau(h.program, X, X, X)
        err = np.abs(d[ok] - r[ok])
        assert err.max() < tau, (name, float(err.max()), tau)


WIDENED = [('pawn', ((-0.8, -0.8, -0.1), (0.8, 0.8, 2.6)), 0.025), ('weave', ((-13, -13, -1), (13, 13, 1)), 0.12),
           ('x_polygon', ((-0.6, -0.3, -0.4), (1.5, 1.8, 0.4)), 0.02), ('slice', ((-1.3, -1.3, -0.4), (1.5, 1.3, 0.4)), 0.02),
           ('extrude_to', ((-1, -1, -1), (1, 1, 1)), 0.02), ('wrap_around', ((-1.6, -1.6, -0.5), (1.6, 1.6, 0.5)), 0.02),
           ('revolve_hex', ((-1.7, -1.7, -0.4), (1.7, 1.7, 0.8)), 0.025)]


@pytest.mark.parametrize('name,bounds,step', WIDENED)
def test_generate_widened_nodes(sdf, ref, name, bounds, step):
    """SURVEY.md section 8(f) rank 1 nodes through the whole pipeline (default precision) vs the oracle."""
    f = models.widened_corpus(sdf)[name]
    pts, info = sdf.generate(f, step=step, bounds=bounds, verbose=False, return_info=True)
    rpts, rstats = ref.generate(f, step=step, bounds=bounds)
    assert (info['skipped'], info['empty'], info['nonempty']) == (rstats['skipped'], rstats['empty'], rstats['nonempty'])
    rep = compare_soups(pts, rpts)
    assert rep['count_equal'] and rep['max_vertex_dist'] <= 1e-5, rep
    assert rep['exact_triangles'] >= 0.999 * rep['n_ref'], rep


def test_generate_auto_bounds_matches_reference_goldens(sdf, golden_dir):
    """samples=N^3 with _estimate_bounds on the GPU evaluator: bounds, step, lattice and batch
    classification equal the real reference's."""
    g = np.load(golden_dir + '/engine.npz')
    from sdf_b200 import core
    for name in ('sphere', 'csg', 'gearlike', 'blobby', 'knurling'):
        f = models.CONFIGS[name](sdf)
        b = core._estimate_bounds(f)
        assert np.array_equal(np.array(b), g['bounds_' + name]), (name, b, g['bounds_' + name])
    for name, n in (('sphere', 64), ('csg', 256)):
        f = models.CONFIGS[name](sdf)
        pts, info = sdf.generate(f, samples=n ** 3, verbose=False, return_info=True)
        gold = g['auto_' + name]
        assert (info['skipped'], info['empty'], info['nonempty'], info['active_cubes']) == tuple(gold[:4]), (name, info)
        assert info['dims'] == tuple(gold[4:7])


def test_skip_flags_match_reference(sdf, golden_dir):
    g = np.load(golden_dir + '/engine.npz')
    from sdf_b200 import core
    for name, n in SMALL:
        f = models.CONFIGS[name](sdf)
        _, L = models.CONFIG_GRIDS[name]
        _, _, (X, Y, Z) = core.make_lattice(f, 2.0 * L / n, ((-L, -L, -L), (L, L, L)))
        h, _ = core.program_handle(f)
        flags = h.skip_flags(X, Y, Z, 32).reshape(-1)
        assert np.array_equal(flags, g['flags_%s_%d' % (name, n)]), name


def test_sparse_off_and_batch_size_invariance(sdf, ref):
    """sparse=False evaluates every batch; the triangle multiset does not depend on batch_size."""
    f = models.cfg_csg(sdf)
    bounds, step = ((-1, -1, -1), (1, 1, 1)), 2.0 / 80
    base, info = sdf.generate(f, step=step, bounds=bounds, verbose=False, sparse=False, precision='f64', return_info=True)
    assert info['skipped'] == 0
    for bs in (8, 16, 24, 40, 64, 80):
        pts, inf = sdf.generate(f, step=step, bounds=bounds, verbose=False, sparse=False, batch_size=bs,
                                precision='f64', return_info=True)
        rep = compare_soups(pts, base)
        # voxel coordinates are batch-local float32, so world positions may differ by an ulp
        assert rep['count_equal'] and rep['max_vertex_dist'] < 1e-6, (bs, rep)
    rpts, rstats = ref.generate(f, step=step, bounds=bounds, sparse=False, batch_size=40)
    pts, inf = sdf.generate(f, step=step, bounds=bounds, verbose=False, sparse=False, batch_size=40,
                            precision='f64', return_info=True)
    assert (inf['skipped'], inf['empty'], inf['nonempty']) == (rstats['skipped'], rstats['empty'], rstats['nonempty'])
    assert compare_soups(pts, rpts)['count_equal']


def test_edge_cases(sdf, ref):
    """Ragged lattices (tail batches of 1..3 planes), single-batch grids, empty results, capacity retry."""
    f = models.cfg_sphere(sdf)
    for n in (33, 34, 65, 66, 31):
        L = 1.25
        step = 2 * L / n
        pts, info = sdf.generate(f, step=step, bounds=((-L, -L, -L), (L, L, L)), verbose=False, precision='f64',
                                 return_info=True)
        rpts, rstats = ref.generate(f, step=step, bounds=((-L, -L, -L), (L, L, L)))
        assert (info['skipped'], info['empty'], info['nonempty']) == (rstats['skipped'], rstats['empty'], rstats['nonempty']), n
        assert compare_soups(pts, rpts)['count_equal'], n
    # no surface inside the bounds: every batch skipped or empty, zero triangles
    pts, info = sdf.generate(f, step=0.05, bounds=((2, 2, 2), (3, 3, 3)), verbose=False, return_info=True)
    assert len(pts) == 0 and info['nonempty'] == 0
    # capacity overflow is reported, never truncated
    from sdf_b200 import core, _abi
    _, _, (X, Y, Z) = core.make_lattice(f, 2.5 / 64, ((-1.25,) * 3, (1.25,) * 3))
    h, _ = core.program_handle(f)
    small = np.empty((10, 9), np.float32)
    with pytest.raises(_abi.CapacityError) as ei:
        h.generate(X, Y, Z, 32, True, 'f64', 0.0, small, 10)
    assert ei.value.needed > 10
    n_tri, stats, _ = h.generate(X, Y, Z, 32, True, 'f64', 0.0, None, 0)      # count only
    assert n_tri == ei.value.needed


def test_save_stl_bytes_equal_the_reference_writer(sdf, ref, tmp_path):
    """f.save('x.stl'): records built on the device are byte-identical to write_binary_stl (sdf/stl.py) applied
    to the same triangles (compared as sets of 50-byte records: batch order is not preserved)."""
    f = models.cfg_csg(sdf)
    bounds, step = ((-1, -1, -1), (1, 1, 1)), 2.0 / 64
    path = tmp_path / 'a.stl'
    f.save(str(path), step=step, bounds=bounds, verbose=False)
    raw = path.read_bytes()
    n = int(np.frombuffer(raw[80:84], '<u4')[0])
    assert raw[:80] == b'\x00' * 80 and len(raw) == 84 + 50 * n
    pts = f.generate(step=step, bounds=bounds, verbose=False)
    assert n * 3 == len(pts)
    host = tmp_path / 'b.stl'
    sdf.write_binary_stl(str(host), pts)
    def sorted_records(buf):
        rows = np.frombuffer(buf, np.uint8).reshape(-1, 50)
        return rows[np.lexsort(rows.T[::-1])]
    assert np.array_equal(sorted_records(raw[84:]), sorted_records(host.read_bytes()[84:]))


def test_user_closure_is_callable_and_meshable_but_never_lowered(sdf):
    """README "writing your own SDFs": a Python closure cannot be lowered to the CUDA evaluator (`ir.lower` raises), yet
    it stays callable, composes with library shapes, and generate() meshes it batch by batch with marching cubes on the
    device (tests/test_gpu_closure.py compares that mesh with the library tree's)."""
    @sdf.sdf3
    def my_sphere(r):
        def f(p):
            return np.linalg.norm(p, axis=1) - r
        return f
    s = my_sphere(1.0)
    assert abs(float(s(np.array([[2.0, 0, 0]]))[0, 0]) - 1.0) < 1e-15
    u = s | sdf.sphere(0.5).translate((1.2, 0, 0))
    from sdf_b200 import ir
    with pytest.raises(sdf.NotLowerableError):
        ir.lower(u)
    assert abs(float(u(np.array([[1.2, 0, 0]]))[0, 0]) + 0.5) < 1e-15          # the library sphere, evaluated on the GPU
    pts, info = u.generate(step=0.1, bounds=((-1.2, -1.2, -1.2), (1.9, 1.2, 1.2)), verbose=False, return_info=True)
    assert info['triangles'] > 500 and info['nonempty'] > 0

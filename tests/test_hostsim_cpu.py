"""The DEVICE CODE itself, executed on the host: sdf_b200/csrc/{sdf_ops.cuh, sdf_kernels.cuh} compiled by g++
against a SIMT stand-in (tests/hostsim/, CUDA threads as fibers) and driven through the same sequence of kernel
launches as sdfb200.cu's generate_impl.  No GPU is involved and nothing of the product is replaced: this is how
the CPU suite -- the only suite the build container can run -- checks the kernels' logic (index arithmetic,
ballot masks, classification, list compaction, table look-ups, the marching-cubes arithmetic, world mapping)
against the oracle and the real-reference goldens.  The tests mirror tests/test_gpu_parity.py, which runs the
same checks on the B200 through the C-ABI.

What it cannot show: anything about speed, memory ordering, or the MUFU units' rounding (the FAST math policy
runs its own formulas on correctly rounded host functions, with a pseudo-random error of the documented size
injected where the test says so).
"""
import numpy as np
import pytest

import models
import util_hostsim as U
from util_cmp import compare_soups, mesh_topology

SMALL = [('sphere', 64), ('csg', 64), ('gearlike', 96), ('blobby', 96), ('knurling', 96)]


@pytest.fixture(scope='module')
def S():
    import sdf_b200
    return sdf_b200


@pytest.fixture(scope='module')
def ref():
    from oracle import sdf_ref
    return sdf_ref


@pytest.fixture(scope='module')
def sims(S):
    """HostSim per model name (the interpreter build: one shared library for every program)."""
    from sdf_b200.ir import lower
    cache = {}

    def get(f):
        key = id(f)
        if key not in cache:
            cache[key] = (f, U.HostSim(lower(f)))
        return cache[key][1]
    return get


def _lattice(S, f, name, n):
    from sdf_b200 import core
    from sdf_b200.ir import lower
    _, L = models.CONFIG_GRIDS[name]
    bounds, step = ((-L, -L, -L), (L, L, L)), 2.0 * L / n
    _, _, (X, Y, Z) = core.make_lattice(f, step, bounds)
    return bounds, step, (X, Y, Z), core.mixed_tau(lower(f), X, Y, Z)


# ------------------------------------------------------------------------------------------ evaluator (a5, node set)

def test_device_node_library_matches_reference_goldens(S, sims, golden_dir):
    """sdf_ops.cuh in float64 (eval_stream_f64_body) against vectors produced by the REAL reference: all 114 models."""
    g = np.load(golden_dir + '/node_eval.npz')
    P = g['points']
    bad = {}
    for name, f in models.build(S).items():
        d = sims(f).eval_f64(P)
        r = g['d_' + name]
        assert np.array_equal(np.isnan(d), np.isnan(r)), name            # numpy's NaN-propagating minimum / maximum
        ok = np.isfinite(r)
        err = np.max(np.abs(d[ok] - r[ok]) / np.maximum(1.0, np.abs(r[ok]))) if ok.any() else 0.0
        if err > 5e-13:
            bad[name] = err
    assert not bad, bad


def test_device_node_library_on_random_trees(S, sims, ref):
    """120 seeded random compositions (the trees tests/test_oracle_cpu.py pins to the real reference)."""
    P = models.random_points(512, seed=11)
    for i, f in enumerate(models.random_trees(S, count=40, seed=424242)):
        d = sims(f).eval_f64(P)
        r = ref.evaluate(f, P).reshape(-1)
        ok = np.isfinite(r)
        assert np.array_equal(np.isnan(d), np.isnan(r)), i
        assert np.all(np.abs(d[ok] - r[ok]) <= 1e-11 * np.maximum(1.0, np.abs(r[ok]))), i


def test_streamed_float32_evaluator_ring_and_ragged_tails(S, sims, ref):
    """eval_stream_f32_body: the 3-stage bulk-copy ring (every stage reused), the ragged last tile, the unaligned
    block that must take the plain-load path, and more tiles than CTAs."""
    f = models.cfg_csg(S)
    hs = sims(f)
    for n, grid in ((1, 1), (3, 2), (1023, 1), (1024, 1), (1025, 2), (4099, 1), (9 * 1024 + 5, 2)):
        P = models.random_points(max(n, 8) + 1, seed=n).astype(np.float32)
        r = ref.evaluate(f, P.astype(np.float64)).reshape(-1)
        for off in (0, 1):                                               # off = 1: base pointer 12 bytes past 16-byte alignment
            d = hs.eval_f32(P[off:off + n], grid=grid)
            assert d.shape == (n,)
            assert np.max(np.abs(d - r[off:off + n])) < 1e-4, (n, off)


def test_float32_policies_against_the_float64_oracle(S, sims, ref):
    """Accurate float32 within 256 ulp of max(1, |p|, |d|); the FAST policy -- with a pseudo-random error of 2 ulp
    injected into every approximate division / square root and 4e-7 into every MUFU sine / cosine -- inside the guard
    band tau, which is what precision `exact` relies on (its float64 pass re-evaluates everything below tau)."""
    from sdf_b200 import core
    from sdf_b200.ir import lower
    P64 = models.random_points(1 << 13, seed=5, L=2.5)
    P32 = P64.astype(np.float32)
    X = np.array([-2.5, 2.5])
    allm = models.build(S)
    for name, f in allm.items():
        hs = sims(f)
        r = ref.evaluate(f, P32.astype(np.float64)).reshape(-1)
        d = hs.eval_f32(P32)
        ok = np.isfinite(r) & np.isfinite(d)
        assert ok.mean() > 0.99, name
        tol = 256 * np.finfo(np.float32).eps * np.maximum(1.0, np.maximum(np.abs(r[ok]), 3.0))
        assert np.all(np.abs(d[ok] - r[ok]) <= tol), name
        hs.set_fast_ulps(2.0)
        try:
            df = hs.eval_f32(P32, fast=True)
        finally:
            hs.set_fast_ulps(0.0)
        okf = np.isfinite(r) & np.isfinite(df) & (np.abs(r) < 1.0)
        tau = core.mixed_tau(lower(f), X, X, X)
        assert np.abs(df[okf] - r[okf]).max() < tau, (name, float(np.abs(df[okf] - r[okf]).max()), tau)


# ------------------------------------------------------------------------------------------ skip pre-pass (a3)

def test_skip_prepass_flags_equal_the_real_reference(S, sims, golden_dir):
    """skip_prepass_body (nine lanes per batch, shuffles + ballot) against flags produced by the REAL core._skip."""
    g = np.load(golden_dir + '/engine.npz')
    from sdf_b200 import core
    for name, n in [('sphere', 64), ('csg', 96), ('gearlike', 128), ('blobby', 128), ('knurling', 160)]:
        f = models.CONFIGS[name](S)
        _, L = models.CONFIG_GRIDS[name]
        _, _, (X, Y, Z) = core.make_lattice(f, 2.0 * L / n, ((-L, -L, -L), (L, L, L)))
        flags = sims(f).skip_flags(X, Y, Z, 32).reshape(-1)
        assert np.array_equal(flags, g['flags_%s_%d' % (name, n)]), name


# ------------------------------------------------------------------------------------------ the tile pipeline (a1-a7)

@pytest.mark.parametrize('name,n', SMALL)
def test_pipeline_exact_is_bit_identical_to_the_oracle(S, sims, ref, name, n):
    """Pre-pass, sign evaluation by autonomous warps, classification, float64 refinement, extraction, batch statistics:
    every triangle of the default precision equals the oracle's, bit for bit; the mesh is watertight."""
    f = models.CONFIGS[name](S)
    bounds, step, (X, Y, Z), tau = _lattice(S, f, name, n)
    tris, st = sims(f).generate(X, Y, Z, 32, True, 'exact', tau)
    rpts, rst = ref.generate(f, step=step, bounds=bounds)
    assert (st['skipped'], st['empty'], st['nonempty']) == (rst['skipped'], rst['empty'], rst['nonempty'])
    rep = compare_soups(tris.reshape(-1, 3), rpts)
    assert rep['count_equal'] and rep['exact_triangles'] == rep['n_ref'] and rep['n_ref'] > 10000, rep
    assert st['refined'] < 0.2 * st['evaluated']
    topo = mesh_topology(tris.reshape(-1, 3))
    assert topo['boundary'] == 0 and topo['nonmanifold'] == 0 and topo['repeated_directed'] == 0, topo


def test_pipeline_classification_equals_reference_goldens(S, sims, golden_dir):
    """Batch classes and active cubes of the REAL reference (tests/golden/engine.npz) at the two sizes shared with it."""
    g = np.load(golden_dir + '/engine.npz')
    for name, n in (('sphere', 64), ('csg', 96)):
        f = models.CONFIGS[name](S)
        _, _, (X, Y, Z), tau = _lattice(S, f, name, n)
        _, st = sims(f).generate(X, Y, Z, 32, True, 'exact', tau)
        gold = g['cls_%s_%d' % (name, n)]
        assert (st['skipped'], st['empty'], st['nonempty'], st['active_cubes']) == tuple(int(v) for v in gold[:4]), (name, st)


@pytest.mark.parametrize('precision', ['f64', 'mixed', 'f32'])
def test_value_tile_kernel_precisions(S, sims, ref, precision):
    """tile_eval_body (the CTA-per-tile kernel of the other precisions): f64 bit-identical, mixed within an ulp, f32
    the same surface within the lattice step (csg: box faces ON lattice planes, so float32 alone misplaces cubes)."""
    name, n = 'csg', 64
    f = models.CONFIGS[name](S)
    bounds, step, (X, Y, Z), tau = _lattice(S, f, name, n)
    tris, st = sims(f).generate(X, Y, Z, 32, True, precision, tau)
    rpts, rst = ref.generate(f, step=step, bounds=bounds)
    rep = compare_soups(tris.reshape(-1, 3), rpts)
    if precision == 'f64':
        assert rep['count_equal'] and rep['exact_triangles'] == rep['n_ref'], rep
        assert st['refined'] == 0
    elif precision == 'mixed':
        assert rep['count_equal'] and rep['max_vertex_dist'] <= 1e-5, rep
        assert 0 < st['refined'] < 0.05 * st['evaluated']
    else:
        assert abs(rep['n_got'] - rep['n_ref']) < 0.01 * rep['n_ref'] and rep['max_vertex_dist'] < 2 * step, rep


def test_guard_band_absorbs_the_fast_policy_error(S, sims):
    """Precision `exact` with 4 ulp of pseudo-random error injected into every approximate operation of the FAST
    pass still equals precision `f64` triangle for triangle, also on trees built to amplify rounding errors
    (tests/test_gpu_guardband.py runs the same trees on the real MUFU units)."""
    trees = {
        'gearlike': models.CONFIGS['gearlike'](S),
        'bend_radial_thin_band': S.box((1.2, 1.2, 0.6)).bend_radial(0.4, 0.401, 0.35),
        'twist_fast': S.box((0.8, 0.8, 1.6)).twist(40.0),
        'elastic_bend': S.box((1.4, 0.5, 0.5)).bend_linear(S.X * -0.7, S.X * 0.7, S.Z * 0.3, S.ease.in_out_elastic),
        'scaled_up': S.sphere(1e-3).scale(800.0) - S.box(0.9),
    }
    from sdf_b200 import core
    from sdf_b200.ir import lower
    for name, f in trees.items():
        L = 2.5 if name == 'gearlike' else 1.2
        X = np.arange(-L, L, 2 * L / 64)
        tau = core.mixed_tau(lower(f), X, X, X)
        hs = sims(f)
        b, sb = hs.generate(X, X, X, 32, True, 'f64', tau)
        hs.set_fast_ulps(4.0)
        try:
            a, sa = hs.generate(X, X, X, 32, True, 'exact', tau)
        finally:
            hs.set_fast_ulps(0.0)
        assert (sa['skipped'], sa['empty'], sa['nonempty']) == (sb['skipped'], sb['empty'], sb['nonempty']), name
        rep = compare_soups(a.reshape(-1, 3), b.reshape(-1, 3))
        assert rep['count_equal'] and rep['exact_triangles'] == rep['n_ref'] and rep['n_ref'] > 0, (name, rep)


def test_guard_band_on_the_model_corpus_and_random_trees(S, sims):
    """The same statement over breadth: every third of the 114 models and of the 120 random compositions on a 33^3
    lattice (offset from the origin so that no sample sits on a symmetry plane), 4 ulp injected: `exact` == `f64`.
    (All 234 trees on a 40^3 lattice were run once by hand: no mismatch.)"""
    from sdf_b200 import core
    from sdf_b200.ir import lower
    trees = list(models.build(S).items())[::3] + [('random%d' % i, f) for i, f in
                                                   enumerate(models.random_trees(S, count=120, seed=424242))][::3]
    X = np.arange(-1.6, 1.6, 3.2 / 33) + 0.0131
    checked = 0
    for name, f in trees:
        prog = lower(f)
        if len(prog.code) > 512:
            continue
        tau = core.mixed_tau(prog, X, X, X)
        hs = sims(f)
        b, _ = hs.generate(X, X, X, 32, False, 'f64', tau, grids=(1, 1, 1, 1))
        hs.set_fast_ulps(4.0)
        try:
            a, _ = hs.generate(X, X, X, 32, False, 'exact', tau, grids=(1, 1, 1, 1))
        finally:
            hs.set_fast_ulps(0.0)
        assert len(a) == len(b) and np.array_equal(canonical(a), canonical(b)), (name, len(a), len(b), tau)
        checked += 1
    assert checked > 70


def test_result_does_not_depend_on_the_launch_shape(S, sims):
    """One CTA per kernel or more CTAs than work items: the same multiset of triangles (work cursors, grid-stride
    loops, last-CTA statistics)."""
    f = models.cfg_csg(S)
    _, _, (X, Y, Z), tau = _lattice(S, f, 'csg', 64)
    hs = sims(f)
    base, sb = hs.generate(X, Y, Z, 32, True, 'exact', tau, grids=(1, 1, 1, 1))
    for grids in ((3, 2, 5, 2), (40, 9, 7, 70)):
        t, st = hs.generate(X, Y, Z, 32, True, 'exact', tau, grids=grids)
        assert st == sb, (grids, st, sb)
        rep = compare_soups(t.reshape(-1, 3), base.reshape(-1, 3))
        assert rep['count_equal'] and rep['exact_triangles'] == rep['n_ref'], (grids, rep)


def test_result_does_not_depend_on_the_thread_schedule(S, sims, ref):
    """The stand-in's racecheck: the threads of a CTA (and the CTAs of a launch) are run in ascending order, in descending
    order and in a fresh random permutation after every rendezvous.  Between two barriers each thread runs to completion, so a read that
    is not ordered after its write by a barrier / warp collective sees the old value under one of the schedules -- and
    the triangles or the statistics change.  All kernels of the step, both tile kernels, the stream evaluator's ring."""
    f = models.cfg_csg(S)
    hs = sims(f)
    _, _, (X, Y, Z), tau = _lattice(S, f, 'csg', 64)
    P = models.random_points(5000, seed=2).astype(np.float32)
    base = {}
    try:
        for mode in (0, 1, 2, 7):
            hs.set_schedule(mode)
            for prec in ('exact', 'mixed'):
                t, st = hs.generate(X, Y, Z, 32, True, prec, tau, grids=(3, 2, 2, 3))
                key = canonical(t)
                if mode == 0:
                    base[prec] = (key, st)
                else:
                    assert st == base[prec][1] and np.array_equal(key, base[prec][0]), (mode, prec)
            d = hs.eval_f32(P, fast=True, grid=1)
            if mode == 0:
                base['eval'] = d
            else:
                assert np.array_equal(d, base['eval']), mode
            fl = hs.skip_flags(X, Y, Z, 32)
            base.setdefault('flags', fl)
            assert np.array_equal(fl, base['flags']), mode
    finally:
        hs.set_schedule(0)


def canonical(tris):
    from util_cmp import canonical_triangles
    return canonical_triangles(tris.reshape(-1, 3))


def test_ragged_lattices_batch_sizes_and_empty_results(S, sims, ref):
    """Tail batches of 1..3 planes, lattices smaller than a batch, batch sizes that are not the tile size
    (sub-tiles: 40, 64), sparse off, nothing to mesh, a result buffer that is too small."""
    f = models.cfg_sphere(S)
    hs = sims(f)
    L = 1.25
    for n, precs in ((31, ('exact',)), (33, ('exact', 'f64')), (34, ('f64',)), (65, ('exact',)), (66, ('f64',))):
        X = np.arange(-L, L, 2 * L / n)
        for prec in precs:
            t, st = hs.generate(X, X, X, 32, True, prec, 1e-5)
            rpts, rst = ref.generate(f, step=2 * L / n, bounds=((-L,) * 3, (L,) * 3))
            assert (st['skipped'], st['empty'], st['nonempty']) == (rst['skipped'], rst['empty'], rst['nonempty']), (n, prec)
            rep = compare_soups(t.reshape(-1, 3), rpts)
            assert rep['count_equal'] and rep['exact_triangles'] == rep['n_ref'], (n, prec, rep)
    g = models.cfg_csg(S)
    hg = sims(g)
    X = np.arange(-1.0, 1.0, 2.0 / 80)
    for bs, sparses in ((16, (True,)), (40, (True, False)), (64, (True,)), (80, (False,))):
        for sparse in sparses:
            t, st = hg.generate(X, X, X, bs, sparse, 'exact', 2e-5)
            rpts, rst = ref.generate(g, step=2.0 / 80, bounds=((-1,) * 3, (1,) * 3), batch_size=bs, sparse=sparse)
            assert (st['skipped'], st['empty'], st['nonempty']) == (rst['skipped'], rst['empty'], rst['nonempty']), (bs, sparse)
            rep = compare_soups(t.reshape(-1, 3), rpts)
            assert rep['count_equal'] and rep['exact_triangles'] == rep['n_ref'], (bs, sparse, rep)
    # no surface inside the bounds: every batch skipped or empty, no triangle
    X = np.arange(2.0, 3.0, 0.05)
    t, st = hs.generate(X, X, X, 32, True, 'exact', 1e-5)
    assert len(t) == 0 and st['nonempty'] == 0 and st['triangles'] == 0
    # capacity overflow: the count is still complete, nothing is written past the buffer
    X = np.arange(-L, L, 2 * L / 64)
    full, st = hs.generate(X, X, X, 32, True, 'exact', 1e-5)
    few, st2 = hs.generate(X, X, X, 32, True, 'exact', 1e-5, cap=100)
    assert st2['triangles'] == st['triangles'] == len(full) and len(few) == 100


# ------------------------------------------------------------------------------------------ run-time specialised code

SPEC = [('gearlike', 96, {}),                                             # the bench tree: runs of 11 along z
        ('knurling', 64, {'SDFB200_JIT_SHARE': 0, 'SDFB200_JIT_N': 6}),   # runs of 6 along x (3 idle slots per line)
        ('csg', 64, {'SDFB200_JIT_SHARE': 1, 'SDFB200_JIT_N': 3})]        # runs of 3 along y


@pytest.mark.parametrize('name,n,env', SPEC)
def test_specialised_evaluator_and_thread_run_shapes(S, ref, name, n, env):
    """The straight-line evaluator sdfb200.cu's jit_source() generates for a tree (taken from the library itself,
    NVRTC needs no GPU), compiled for the host: every thread-run shape of tile_sign_body -- share axis x / y / z,
    11 / 6 / 3 samples per run, the classification kernel instantiated for the same axis -- gives the oracle's mesh."""
    from sdf_b200.ir import lower
    f = models.CONFIGS[name](S)
    prog = lower(f)
    try:
        spec = U.spec_eval_text(prog, kernel=4, env=env)
    except Exception as e:                                                # no libnvrtc on this machine
        pytest.skip('specialised source unavailable: %s' % e)
    hs = U.HostSim(prog, spec)
    is_spec, kn, share = hs.shape()
    assert is_spec and (not env or (share, kn) == (env['SDFB200_JIT_SHARE'], env['SDFB200_JIT_N'])), (kn, share)
    bounds, step, (X, Y, Z), tau = _lattice(S, f, name, n)
    tris, st = hs.generate(X, Y, Z, 32, True, 'exact', tau)
    rpts, rst = ref.generate(f, step=step, bounds=bounds)
    assert (st['skipped'], st['empty'], st['nonempty']) == (rst['skipped'], rst['empty'], rst['nonempty'])
    rep = compare_soups(tris.reshape(-1, 3), rpts)
    assert rep['count_equal'] and rep['exact_triangles'] == rep['n_ref'] and rep['n_ref'] > 5000, rep
    P = models.random_points(2048, seed=3)
    r = ref.evaluate(f, P).reshape(-1)
    assert np.max(np.abs(hs.eval_f64(P) - r)) < 1e-12


# ------------------------------------------------------------------------------------------ marching cubes (a6, a7)

def test_extraction_kernel_equals_the_c_restatement_on_every_configuration(S, sims, ref):
    """extract_body against oracle/mc_lewiner.c: 16384 cubes built from all 256 sign patterns x 64 magnitude sets
    (every face / interior decision of the ambiguous cases both ways), and 33^3 / ragged noise volumes."""
    hs = sims(models.cfg_sphere(S))
    rng = np.random.default_rng(20260923)
    ax = np.arange(32.0)
    for rep_i in range(4):
        # 16 x 16 x 16 test cubes side by side in one 32^3 volume (the cubes in between mix their corners: more cases)
        vol = np.empty((32, 32, 32), np.float32)
        index = rng.permutation(4096) % 256
        for c, idx in enumerate(index):
            i, j, k = c >> 8, (c >> 4) & 15, c & 15
            sign = np.array([1.0 if (idx >> q) & 1 else -1.0 for q in range(8)])
            mag = rng.uniform(0.02, 1.0, 8) if (c + rep_i) % 4 else rng.choice([1e-3, 0.5, 1.0, 30.0], 8)
            for q in range(8):                       # Lewiner corner q -> array index (Lewiner x = last array axis)
                lx, ly, lz = (q ^ (q >> 1)) & 1, (q >> 1) & 1, (q >> 2) & 1
                vol[2 * i + lz, 2 * j + ly, 2 * k + lx] = sign[q] * mag[q]
        got = hs.mc_volume(vol, ax, ax, ax, grid=2 + rep_i)
        want = ref.marching_cubes_soup(vol, 0.0)
        r = compare_soups(got.reshape(-1, 3), want)
        assert r['count_equal'] and r['exact_triangles'] == r['n_ref'] and r['n_ref'] > 50000, (rep_i, r)
    for shape in ((33, 33, 33), (7, 33, 12), (2, 2, 33), (33, 2, 5)):
        vol = rng.normal(size=shape).astype(np.float32)
        axes = [np.arange(float(s)) * 0.25 - 1.0 for s in shape]          # exact in float: step 1/4
        got = hs.mc_volume(vol, *axes, grid=3)
        want = ref.marching_cubes_soup(vol, 0.0).astype(np.float64) * 0.25 - 1.0
        r = compare_soups(got.reshape(-1, 3), want)
        assert r['count_equal'] and r['exact_triangles'] == r['n_ref'] and r['n_ref'] > 0, (shape, r)
    # NaN, infinite, zero, denormal samples (a user's SDF can produce them: ellipsoid's centre is 0/0): the same bits as
    # the C restatement -- a vertex next to a NaN corner is NaN in all three coordinates, as skimage's generic form gives
    from util_cmp import soup_multiset_bits, special_value_volumes
    ax12 = np.arange(12.0)
    for vol in special_value_volumes():
        got = hs.mc_volume(vol, ax12, ax12, ax12, grid=2)
        want = ref.marching_cubes_soup(vol, 0.0)
        assert soup_multiset_bits(got) == soup_multiset_bits(want) and len(got) > 3000
    # degenerate inputs the reference maps to an empty batch (core.py:53-56)
    assert len(hs.mc_volume(np.ones((5, 5, 5), np.float32), *[np.arange(5.0)] * 3)) == 0
    assert len(hs.mc_volume(np.ones((1, 5, 5), np.float32), np.arange(1.0), np.arange(5.0), np.arange(5.0))) == 0


def test_stl_record_kernel_bytes_equal_the_reference_writer(S, sims, tmp_path):
    """stl_records_body: 50-byte records byte-identical to write_binary_stl (sdf/stl.py) on the same triangles."""
    f = models.cfg_csg(S)
    hs = sims(f)
    _, _, (X, Y, Z), tau = _lattice(S, f, 'csg', 64)
    tris, _ = hs.generate(X, Y, Z, 32, True, 'exact', tau)
    for n in (1, 127, 128, 129, len(tris)):
        rec = hs.stl_records(tris[:n])
        path = tmp_path / ('a%d.stl' % n)
        S.write_binary_stl(str(path), tris[:n].reshape(-1, 3))
        raw = path.read_bytes()
        assert int(np.frombuffer(raw[80:84], '<u4')[0]) == n
        assert rec.tobytes() == raw[84:], n


# ------------------------------------------------------------------------------------------ memcheck

def test_kernels_are_clean_under_address_and_undefined_behaviour_sanitizers():
    """tests/hostsim/sanitize_run.py: every kernel body (all four precisions, ragged lattices, sub-tiled batches, the
    stream evaluators' ring and fallback path, a too-small result buffer, caller volumes, STL records) with ASan +
    UBSan instrumentation: no out-of-bounds access to a shared array, a work list or a staged table, no undefined shift."""
    import os
    import subprocess
    import sys
    pre = U.sanitizer_preload()
    if pre is None:
        pytest.skip('no libasan / libubsan next to g++')
    env = dict(os.environ, LD_PRELOAD=pre, ASAN_OPTIONS='detect_leaks=0:detect_stack_use_after_return=0',
               UBSAN_OPTIONS='print_stacktrace=1')
    r = subprocess.run([sys.executable, os.path.join(U.HS, 'sanitize_run.py')], env=env, capture_output=True, text=True,
                       timeout=1500)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0 and 'SANITIZED RUN COMPLETE' in r.stdout, tail
    assert 'ERROR: AddressSanitizer' not in r.stderr and 'runtime error:' not in r.stderr, tail

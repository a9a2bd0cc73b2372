"""Differential sweeps of the device code against the oracle on the HOST execution (tests/hostsim/) -- the exploratory
runs DESIGN.md section 2.1 cites.  No GPU involved; minutes each.

  python tools/hostsim_sweeps.py special-points     evaluator (float64) on 14^3 special points x 114 models + 120 random trees
  python tools/hostsim_sweeps.py odd-lattices       the host-compiled LIBRARY on 412 odd lattices (batch sizes 1..200, one-sample axes)
  python tools/hostsim_sweeps.py guard-band         `exact` (4 ulp injected into the FAST pass) == `f64` on all 234 trees, 40^3
  python tools/hostsim_sweeps.py special-volumes    extraction kernel vs the C restatement on NaN / inf / denormal volumes
  python tools/hostsim_sweeps.py magnitudes         the host-compiled LIBRARY on anisotropic steps, lattices 10 .. 3e6 away from the
                                                    origin, models scaled by 1e-4 .. 1e6, auto bounds (the guard band scales with |coords|)
"""
import itertools
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
warnings.simplefilter('ignore')


def _trees(S, models):
    return list(models.build(S).items()) + [('random%d' % i, f) for i, f in enumerate(models.random_trees(S, count=120, seed=424242))]


def special_points():
    import models
    import sdf_b200 as S
    import util_hostsim as U
    from oracle import sdf_ref
    from sdf_b200.ir import lower
    vals = [-2, -1, -0.5, -0.25, 0.0, -0.0, 0.25, 0.5, 1, 2, 1e-300, -1e-300, 1e-9, 0.3]
    P = np.array(list(itertools.product(vals, repeat=3)), np.float64)
    bad = 0
    trees = _trees(S, models)
    for name, f in trees:
        prog = lower(f)
        if len(prog.code) > 512:
            continue
        d = U.HostSim(prog).eval_f64(P)
        r = sdf_ref.evaluate(f, P).reshape(-1)
        ok = np.isfinite(r) & np.isfinite(d)
        diff = (np.isnan(d) != np.isnan(r)) | (np.isinf(d) != np.isinf(r))
        diff[ok] |= np.abs(d[ok] - r[ok]) > 1e-11 * np.maximum(1.0, np.abs(r[ok]))
        if diff.any():
            bad += 1
            k = int(np.flatnonzero(diff)[0])
            print('DIFF', name, int(diff.sum()), P[k].tolist(), float(d[k]), float(r[k]))
    print('special points: %d points x %d trees, %d trees with differences' % (len(P), len(trees), bad))
    return bad


def odd_lattices():
    import util_hostsim as U
    from sdf_b200 import _abi
    _abi.LIB_PATH = U.build_library()
    os.environ['SDFB200_JIT'] = '0'
    import models
    import sdf_b200 as S
    from oracle import sdf_ref
    from sdf_b200 import core
    from util_cmp import compare_soups
    f = models.cfg_csg(S)

    def oracle(X, Y, Z, bs, sparse):
        pts, sk, em, ne = [], 0, 0, 0
        for job in sdf_ref.batches_of(X, Y, Z, bs):
            res = sdf_ref.worker(f, job, sparse)
            if res is None:
                sk += 1
            elif len(res) == 0:
                em += 1
            else:
                ne += 1
                pts.append(res)
        return (np.concatenate(pts) if pts else np.zeros((0, 3))), (sk, em, ne)
    bad = cases = 0
    for bs in (1, 2, 3, 5, 16, 31, 32, 33, 47, 64, 65, 96, 128, 200):
        for dims in ((40, 40, 40), (1, 40, 40), (40, 1, 40), (40, 40, 1), (2, 2, 2), (33, 65, 34), (97, 5, 3), (3, 70, 129)):
            if bs < 4 and max(dims) > 40:
                continue
            if bs == 1 and dims[0] * dims[1] * dims[2] > 20000:
                dims = tuple(min(d, 12) for d in dims)
            X, Y, Z = (-1.0 + (2.0 / 40) * np.arange(n) for n in dims)
            for sparse in (True, False):
                rpts, rst = oracle(X, Y, Z, bs, sparse)
                for prec in ('exact', 'f64'):
                    pts, info = core.run_batches(f, X, Y, Z, bs, sparse, prec)
                    rep = compare_soups(np.asarray(pts).reshape(-1, 3), rpts)
                    cases += 1
                    if not (rep['count_equal'] and rep['exact_triangles'] == rep['n_ref'] and
                            (info['skipped'], info['empty'], info['nonempty']) == rst):
                        bad += 1
                        print('DIFF', bs, dims, sparse, prec, rep, info, rst)
    print('odd lattices: %d cases, %d differences' % (cases, bad))
    return bad


def guard_band():
    import models
    import sdf_b200 as S
    import util_hostsim as U
    from sdf_b200 import core
    from sdf_b200.ir import lower
    from util_cmp import canonical_triangles
    bad = n = 0
    for name, f in _trees(S, models):
        prog = lower(f)
        if len(prog.code) > 512:
            continue
        hs = U.HostSim(prog)
        X = np.arange(-1.6, 1.6, 3.2 / 40) + 0.0131
        tau = core.mixed_tau(prog, X, X, X)
        b, _ = hs.generate(X, X, X, 32, False, 'f64', tau)
        hs.set_fast_ulps(4.0)
        a, _ = hs.generate(X, X, X, 32, False, 'exact', tau)
        hs.set_fast_ulps(0.0)
        n += 1
        if len(a) != len(b) or not np.array_equal(canonical_triangles(a.reshape(-1, 3)), canonical_triangles(b.reshape(-1, 3))):
            bad += 1
            print('DIFF', name, len(a), len(b), tau)
    print('guard band: %d trees, %d differences' % (n, bad))
    return bad


def special_volumes():
    import models
    import sdf_b200 as S
    import util_hostsim as U
    from oracle import sdf_ref
    from sdf_b200.ir import lower
    from util_cmp import soup_multiset_bits, special_value_volumes
    hs = U.HostSim(lower(models.cfg_sphere(S)))
    ax = np.arange(12.0)
    bad = 0
    vols = special_value_volumes(count=40)
    for i, vol in enumerate(vols):
        if soup_multiset_bits(hs.mc_volume(vol, ax, ax, ax)) != soup_multiset_bits(sdf_ref.marching_cubes_soup(vol, 0.0)):
            bad += 1
            print('DIFF volume', i)
    print('special volumes: %d volumes, %d differences' % (len(vols), bad))
    return bad


def magnitudes():
    import util_hostsim as U
    from sdf_b200 import _abi
    _abi.LIB_PATH = U.build_library()
    os.environ['SDFB200_JIT'] = '0'
    import models
    import sdf_b200 as S
    from oracle import sdf_ref
    from util_cmp import compare_soups
    state = dict(bad=0, cases=0)

    def check(f, label, **kw):
        pts, info = S.generate(f, verbose=False, return_info=True, **kw)
        rpts, rst = sdf_ref.generate(f, **kw)
        rep = compare_soups(pts, rpts)
        state['cases'] += 1
        if not (rep['count_equal'] and rep['exact_triangles'] == rep['n_ref'] and rep['n_ref'] > 0 and
                (info['skipped'], info['empty'], info['nonempty']) == (rst['skipped'], rst['empty'], rst['nonempty'])):
            state['bad'] += 1
            print('DIFF', label, rep, info, rst)
    g = models.cfg_gearlike(S)
    for step in ((0.05, 0.07, 0.031), (0.11, 0.04, 0.04), (0.033, 0.033, 0.2)):
        check(g, 'anisotropic %r' % (step,), step=step, bounds=((-2.2, -2.3, -0.9), (2.25, 2.2, 0.95)))
    for off in (10.0, 1e3, 1e5, -3e6):
        f = S.sphere(1).translate((off, 0.5 * off, 0)) - S.box(0.8).translate((off + 0.6, 0.5 * off, 0.2)).k(0.05)
        check(f, 'offset %g' % off, step=2.6 / 48, bounds=((off - 1.3, 0.5 * off - 1.3, -1.3), (off + 1.3, 0.5 * off + 1.3, 1.3)))
    for sc in (1e-4, 1e-2, 1e3, 1e6):
        f = (S.sphere(1) & S.box(1.5)).scale(sc) | S.capsule((-sc, 0, 0), (sc, 0, 0), 0.3 * sc).k(0.1 * sc)
        check(f, 'scale %g' % sc, step=2.6 * sc / 48, bounds=((-1.3 * sc,) * 3, (1.3 * sc,) * 3))
    for name in ('sphere', 'csg'):
        check(models.CONFIGS[name](S), 'samples %s' % name, samples=40 ** 3)
    print('magnitudes: %d cases, %d differences' % (state['cases'], state['bad']))
    return state['bad']


if __name__ == '__main__':
    sweeps = {'magnitudes': magnitudes, 'special-points': special_points, 'odd-lattices': odd_lattices, 'guard-band': guard_band,
              'special-volumes': special_volumes}
    if len(sys.argv) != 2 or sys.argv[1] not in sweeps:
        sys.exit(__doc__)
    sys.exit(1 if sweeps[sys.argv[1]]() else 0)

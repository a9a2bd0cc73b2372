"""The lowering (sdf_b200/ir.py) checked on the CPU: every model of the corpus is flattened to the
evaluator's linear program, the program is run by a numpy simulator of the device's register machine
(tests/prog_sim.py) and compared with the oracle's direct evaluation of the tree.  Catches wrong stack
depths, point-slot reuse, fold order and constant layout without a GPU."""
import numpy as np

import models
import prog_sim
import sdf_b200
from oracle import sdf_ref
from sdf_b200 import ir


def test_lowered_programs_evaluate_like_the_tree():
    P = models.random_points(3000, seed=5)
    worst = {}
    for name, f in models.build(sdf_b200).items():
        prog = ir.lower(f)
        got = prog_sim.run(prog, P)
        ref = sdf_ref.evaluate(f, P).reshape(-1)
        assert np.array_equal(np.isnan(got), np.isnan(ref)), name
        ok = np.isfinite(ref)
        err = np.max(np.abs(got[ok] - ref[ok]) / np.maximum(1.0, np.abs(ref[ok]))) if ok.any() else 0.0
        worst[name] = err
    bad = {k: v for k, v in worst.items() if v > 1e-11}
    assert not bad, bad


def test_late_k_and_shared_subtrees_lower_correctly():
    """`.k()` set after the fold was built (dn.py:12 reads it at call time) and a sub-tree used twice."""
    S = sdf_b200
    P = models.random_points(2000, seed=6)
    a, b = S.sphere(1), S.box(1.2)
    f = a | b
    b.k(0.3)                                   # late: must still smooth the union
    c = S.cylinder(0.4).orient(S.X)
    g = (f - c) | c.translate((0, 0, 1.5)) | (c & a)
    for tree in (f, g):
        got = prog_sim.run(ir.lower(tree), P)
        ref = sdf_ref.evaluate(tree, P).reshape(-1)
        assert np.max(np.abs(got - ref)) < 1e-12


def test_random_trees_lower_correctly():
    """300 seeded random trees (depth <= 4) through lower() + simulator vs the oracle's tree walk."""
    rng = np.random.default_rng(20260922)
    P = models.random_points(600, seed=8)
    deepest = 0
    for i in range(300):
        f = models.random_tree(sdf_b200, rng, int(rng.integers(1, 5)))
        try:
            prog = ir.lower(f)
        except ir.NotLowerableError:
            continue                                     # stack / slot budget exceeded: refused, not mis-lowered
        deepest = max(deepest, len(prog.code))
        got = prog_sim.run(prog, P)
        ref = sdf_ref.evaluate(f, P).reshape(-1)
        assert np.array_equal(np.isnan(got), np.isnan(ref)), i
        ok = np.isfinite(ref)
        err = np.max(np.abs(got[ok] - ref[ok]) / np.maximum(1.0, np.abs(ref[ok]))) if ok.any() else 0.0
        assert err < 1e-10, (i, err)
    assert deepest > 100


def test_random_tree_goldens_through_the_simulator(golden_dir):
    """The trees of tests/golden/random_trees.npz (values from the REAL reference): lower() + simulator."""
    g = np.load(golden_dir + '/random_trees.npz')
    P = g['points']
    for i, f in enumerate(models.random_trees(sdf_b200)):
        got = prog_sim.run(ir.lower(f), P)
        r = g['d_%03d' % i]
        assert np.array_equal(np.isnan(got), np.isnan(r)), i
        ok = np.isfinite(r)
        assert np.max(np.abs(got[ok] - r[ok]) / np.maximum(1.0, np.abs(r[ok]))) < 1e-12, i


def test_random_trees_specialise_with_nvrtc():
    """The generated straight-line evaluator of random trees compiles for sm_100a (no GPU needed), in the
    float64 streamed kernel and in the fused tile kernel with whatever sample grouping the plan picks."""
    from sdf_b200 import _abi
    rng = np.random.default_rng(7)
    done = 0
    while done < 8:
        f = models.random_tree(sdf_b200, rng, 4)
        try:
            prog = ir.lower(f)
        except ir.NotLowerableError:
            continue
        h = _abi.ProgramHandle(prog)
        assert h.jit_probe(7) > 10000 and h.jit_probe(4) > 10000
        axis, group, frac = h.jit_plan(3)
        assert axis in (-1, 0, 1, 2) and group in (4, 6, 11) and 0.0 <= frac <= 1.0
        done += 1

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


def _random_tree(rng, depth):
    S = sdf_b200
    u = lambda lo, hi: float(rng.uniform(lo, hi))
    v3 = lambda lo, hi: tuple(float(t) for t in rng.uniform(lo, hi, 3))
    if depth == 0 or rng.random() < 0.15:
        k = rng.integers(0, 12)
        if k == 0: return S.sphere(u(0.3, 1.2), v3(-0.5, 0.5))
        if k == 1: return S.box(v3(0.4, 1.5))
        if k == 2: return S.rounded_box(v3(0.6, 1.5), u(0.05, 0.2))
        if k == 3: return S.torus(u(0.6, 1.2), u(0.1, 0.3))
        if k == 4: return S.capsule(v3(-1, 0), v3(0, 1), u(0.1, 0.4))
        if k == 5: return S.cylinder(u(0.2, 0.8))
        if k == 6: return S.capped_cylinder(v3(-1, 0), v3(0, 1), u(0.2, 0.5))
        if k == 7: return S.ellipsoid(v3(0.4, 1.3))
        if k == 8: return S.hexagon(u(0.4, 1.0)).extrude(u(0.3, 1.5))
        if k == 9: return S.circle(u(0.2, 0.6)).translate((u(0.7, 1.2), 0)).revolve()
        if k == 10: return S.rectangle((u(0.5, 1.2), u(0.3, 0.9))).extrude_to(S.circle(u(0.3, 0.8)), u(0.5, 1.5))
        return S.octahedron(u(0.5, 1.2))
    k = rng.integers(0, 16)
    a = _random_tree(rng, depth - 1)
    if k == 0: return a.translate(v3(-1, 1))
    if k == 1: return a.scale(u(0.5, 1.8))
    if k == 2: return a.rotate(u(-3, 3), v3(-1, 1))
    if k == 3: return a.orient(v3(-1, 1))
    if k == 4: return a.twist(u(-1, 1))
    if k == 5: return a.elongate(v3(0, 0.5))
    if k == 6: return a.circular_array(int(rng.integers(2, 9)), u(0, 1.5))
    if k == 7: return a.repeat(v3(2.5, 4.0), count=(1, 1, 0), padding=int(rng.integers(0, 2)))
    if k == 8: return a.shell(u(0.05, 0.2))
    if k == 9: return a.dilate(u(0.02, 0.2))
    if k == 10: return a.bend_linear(v3(-1, 0), v3(0, 1), v3(-0.3, 0.3))
    b = _random_tree(rng, depth - 1)
    kk = None if rng.random() < 0.5 else u(0.05, 0.4)
    if k == 11: return a.union(b, _random_tree(rng, depth - 1), k=kk)
    if k == 12: return a.difference(b, k=kk)
    if k == 13: return a.intersection(b, k=kk)
    if k == 14: return a.blend(b, k=u(0.1, 0.9))
    return a.transition_linear(b, v3(-1, 0), v3(0, 1))


def test_random_trees_lower_correctly():
    """300 seeded random trees (depth <= 4) through lower() + simulator vs the oracle's tree walk."""
    rng = np.random.default_rng(20260922)
    P = models.random_points(600, seed=8)
    deepest = 0
    for i in range(300):
        f = _random_tree(rng, int(rng.integers(1, 5)))
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


def test_random_trees_specialise_with_nvrtc():
    """The generated straight-line evaluator of random trees compiles for sm_100a (no GPU needed), in the
    float64 streamed kernel and in the fused tile kernel with whatever sample grouping the plan picks."""
    from sdf_b200 import _abi
    rng = np.random.default_rng(7)
    done = 0
    while done < 8:
        f = _random_tree(rng, 4)
        try:
            prog = ir.lower(f)
        except ir.NotLowerableError:
            continue
        h = _abi.ProgramHandle(prog)
        assert h.jit_probe(7) > 10000 and h.jit_probe(4) > 10000
        axis, group, frac = h.jit_plan(3)
        assert axis in (-1, 0, 1, 2) and group in (4, 6, 11) and 0.0 <= frac <= 1.0
        done += 1

#!/usr/bin/env python
"""Generates tests/golden/*.npz by importing the REAL reference (/root/reference/sdf).

Run in the build container only (the GPU box has no /root/reference).  scikit-image is not
installed, so a stub ``skimage.measure`` is inserted before ``import sdf``; every vector here
therefore pins the half of the hot path that does not go through skimage: node evaluation
(d3.py / dn.py closures), _estimate_bounds, the np.arange lattice, _skip and the
skipped / empty / nonempty classification ("nonempty" = a non-skipped batch whose float32
volume has a cube with mixed `v > 0` corner signs, i.e. skimage would emit a vertex).
"""
import itertools
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import models  # noqa: E402


def import_reference():
    stub = types.ModuleType('skimage')
    stub.measure = types.ModuleType('skimage.measure')

    def marching_cubes(volume, level=0):
        raise RuntimeError('skimage is not installed')
    stub.measure.marching_cubes = marching_cubes
    sys.modules['skimage'] = stub
    sys.modules['skimage.measure'] = stub.measure
    sys.path.insert(0, '/root/reference')
    import sdf
    return sdf


def classify(ref, f, bounds, step, batch_size=32, sparse=True):
    core = ref.core
    (x0, y0, z0), (x1, y1, z1) = bounds
    X, Y, Z = np.arange(x0, x1, step), np.arange(y0, y1, step), np.arange(z0, z1, step)
    s = batch_size
    Xs = [X[i:i + s + 1] for i in range(0, len(X), s)]
    Ys = [Y[i:i + s + 1] for i in range(0, len(Y), s)]
    Zs = [Z[i:i + s + 1] for i in range(0, len(Z), s)]
    skipped = empty = nonempty = active = 0
    flags = []
    for job in itertools.product(Xs, Ys, Zs):
        if sparse and core._skip(f, job):
            skipped += 1
            flags.append(1)
            continue
        flags.append(0)
        P = core._cartesian_product(*job)
        vol = f(P).reshape([len(a) for a in job]).astype(np.float32)
        if min(vol.shape) < 2:
            empty += 1
            continue
        pos = vol > 0
        c = pos[:-1, :-1, :-1].astype(np.int8)
        for dx, dy, dz in itertools.product((0, 1), repeat=3):
            if (dx, dy, dz) != (0, 0, 0):
                c = c + pos[dx:pos.shape[0] - 1 + dx, dy:pos.shape[1] - 1 + dy, dz:pos.shape[2] - 1 + dz]
        n = int(np.count_nonzero((c > 0) & (c < 8)))
        active += n
        if n:
            nonempty += 1
        else:
            empty += 1
    return dict(skipped=skipped, empty=empty, nonempty=nonempty, active=active,
                dims=(len(X), len(Y), len(Z)), flags=np.array(flags, np.uint8))


def natural_grids(ref, ref_models):
    """`samples = N^3` with the reference's own bounds estimate.
    Slow (minutes: the real reference evaluates every non-skipped batch): `make_golden.py --natural`."""
    out = {}
    for name, n in (('gearlike', 512), ('blobby', 768), ('knurling', 1024)):
        f = ref_models['cfg_' + name]
        lo, hi = (np.array(v, np.float64) for v in ref.core._estimate_bounds(f))
        step = (float(np.prod(hi - lo)) / n ** 3) ** (1 / 3)
        r = classify(ref, f, (tuple(lo), tuple(hi)), step)
        out['auto_%s' % name] = np.array([r['skipped'], r['empty'], r['nonempty'], r['active'], *r['dims']], np.int64)
        out['autostep_%s' % name] = np.array([step])
        print('auto', name, r['skipped'], r['empty'], r['nonempty'], r['active'], r['dims'], step, flush=True)
        np.savez_compressed(os.path.join(HERE, 'natural.npz'), **out)


def main():
    ref = import_reference()
    ref_models = models.build(ref)
    if '--natural' in sys.argv:
        return natural_grids(ref, ref_models)
    P = models.random_points(4096, seed=1234)
    out = {'points': P}
    for name, f in ref_models.items():
        out['d_' + name] = f(P).reshape(-1)
    np.savez_compressed(os.path.join(HERE, 'node_eval.npz'), **out)
    print('node_eval.npz: %d models x %d points' % (len(ref_models), len(P)))

    eng = {}
    # auto bounds for every config (reference _estimate_bounds)
    for name in models.CONFIGS:
        f = ref_models['cfg_' + name]
        (lo, hi) = ref.core._estimate_bounds(f)
        eng['bounds_' + name] = np.array([lo, hi], np.float64)
    # batch classification on small lattices (same trees, reduced N so it runs in seconds)
    for name, n in (('sphere', 64), ('csg', 96), ('gearlike', 128), ('blobby', 128), ('knurling', 160)):
        f = ref_models['cfg_' + name]
        _, L = models.CONFIG_GRIDS[name]
        r = classify(ref, f, ((-L, -L, -L), (L, L, L)), 2.0 * L / n)
        eng['cls_%s_%d' % (name, n)] = np.array([r['skipped'], r['empty'], r['nonempty'], r['active'], *r['dims']], np.int64)
        eng['flags_%s_%d' % (name, n)] = r['flags']
        print(name, n, r['skipped'], r['empty'], r['nonempty'], r['active'], r['dims'])
    # the survey's natural-grid goldens for configs 1 and 2 (samples = N^3, auto bounds)
    for name, n in (('sphere', 64), ('csg', 256)):
        f = ref_models['cfg_' + name]
        lo, hi = eng['bounds_' + name]
        vol = float(np.prod(hi - lo))
        step = (vol / n ** 3) ** (1 / 3)
        r = classify(ref, f, (tuple(lo), tuple(hi)), step)
        eng['auto_%s' % name] = np.array([r['skipped'], r['empty'], r['nonempty'], r['active'], *r['dims']], np.int64)
        eng['autostep_%s' % name] = np.array([step])
        print('auto', name, r['skipped'], r['empty'], r['nonempty'], r['active'], r['dims'], step)
    # sample_slice (core.py:202-232): values, extent and axis names of the three orientations
    f = ref_models['cfg_csg']
    sb = ((-1.0, -0.9, -0.8), (1.1, 1.0, 0.7))
    for key, kw in (('x', dict(x=0.13)), ('y', dict(y=-0.2)), ('z', dict(z=0.31))):
        a, extent, axes = ref.core.sample_slice(f, w=24, h=17, bounds=sb, **kw)
        eng['slice_' + key] = a
        eng['slice_extent_' + key] = np.array(extent, np.float64)
        eng['slice_axes_' + key] = np.array([ord(c) for c in axes], np.int64)
    # write_binary_stl (stl.py:4-24) on a seeded soup with one degenerate triangle (0/0 normal)
    import tempfile
    rs = np.random.default_rng(77)
    soup = rs.uniform(-3, 3, (60 * 3, 3))
    soup[30:33] = soup[30]                                     # zero-area triangle
    with tempfile.TemporaryDirectory() as td:
        with np.errstate(all='ignore'):
            ref.write_binary_stl(os.path.join(td, 'a.stl'), soup)
        eng['stl_points'] = soup
        eng['stl_bytes'] = np.frombuffer(open(os.path.join(td, 'a.stl'), 'rb').read(), np.uint8)
    np.savez_compressed(os.path.join(HERE, 'engine.npz'), **eng)

    # random compositions (tests/models.py: random_trees) evaluated by the real reference
    Pr = models.random_points(512, seed=4321)
    rt = {'points': Pr}
    for i, f in enumerate(models.random_trees(ref)):
        rt['d_%03d' % i] = f(Pr).reshape(-1)
    np.savez_compressed(os.path.join(HERE, 'random_trees.npz'), **rt)
    print('random_trees.npz: %d trees x %d points' % (len(rt) - 1, len(Pr)))


if __name__ == '__main__':
    main()

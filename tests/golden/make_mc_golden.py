#!/usr/bin/env python
"""Pins the Lewiner marching-cubes half of the path against REAL scikit-image -- wherever it can be imported.

    python tests/golden/make_mc_golden.py            # writes tests/golden/mc_skimage.npz

The reference calls ``skimage.measure.marching_cubes(volume, 0)`` (core.py:16-18, setup.py:14) and nothing in
this repository has been compared with an skimage output, because scikit-image is neither installed nor in the
offline wheelhouse of the build container and of the GPU boxes ("parity unpinned").
This script is the ready-to-run pin: on any machine where ``import skimage`` works it stores, for

  * all 256 corner sign patterns x 64 random magnitude sets (2x2x2 volumes; the ambiguous cases 3, 4, 6, 7, 10,
    12, 13 meet every face / interior decision this way),
  * 8 seeded 33^3 noise volumes (every case, shared edges and faces between cubes),
  * the first non-empty batch (33^3, float32 as skimage sees it) of each of the five BASELINE configurations,
    evaluated by the CPU oracle (numpy float64, bit-identical to the reference's closures),

the volume and ``verts[faces].reshape(-1, 3)`` exactly as core.py:17-18 builds it.  tests/test_mc_skimage.py
then compares the C restatement (oracle/mc_lewiner.c) -- and through it the CUDA extraction kernel, which is
bit-identical to the restatement -- with these soups: it FAILS on a mismatch when the file exists and SKIPS,
loudly, when it does not.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))


def soup_of(measure, volume):
    """core.py:16-18; skimage's ValueError / RuntimeError become an empty soup like core.py:53-56."""
    try:
        verts, faces, _, _ = measure.marching_cubes(volume, 0)
    except (ValueError, RuntimeError):
        return np.zeros((0, 3), np.float32)
    return np.asarray(verts[faces].reshape((-1, 3)), np.float32)


def cube_volumes():
    """(256 * 64, 2, 2, 2) float32: every sign pattern, magnitudes log-uniform in [1e-3, 1] (seeded)."""
    rng = np.random.default_rng(20260923)
    vols = np.empty((256, 64, 2, 2, 2), np.float32)
    for index in range(256):
        mag = np.exp(rng.uniform(np.log(1e-3), 0.0, size=(64, 8))).astype(np.float32)
        sign = np.array([1.0 if (index >> k) & 1 else -1.0 for k in range(8)], np.float32)
        vols[index] = (mag * sign).reshape(64, 2, 2, 2)
    return vols.reshape(-1, 2, 2, 2)


def noise_volumes():
    rng = np.random.default_rng(33)
    return [rng.standard_normal((33, 33, 33)).astype(np.float32) for _ in range(8)]


def config_volumes():
    import models
    import sdf_b200
    from oracle import sdf_ref
    out = {}
    for name, builder in models.CONFIGS.items():
        f = builder(sdf_b200)
        n, L = models.CONFIG_GRIDS[name]
        X, Y, Z = sdf_ref.lattice(((-L, -L, -L), (L, L, L)), 2.0 * L / n)
        for job in sdf_ref.batches_of(X, Y, Z, 32):
            if min(len(a) for a in job) < 2 or sdf_ref.skip(f, job):
                continue
            vol = sdf_ref.evaluate(f, sdf_ref.cartesian_product(*job)).reshape([len(a) for a in job]).astype(np.float32)
            if (vol > 0).any() and (vol <= 0).any():
                out[name] = vol
                break
    return out


def main():
    try:
        import skimage
        from skimage import measure
    except ImportError as e:
        print('scikit-image is not importable here (%s): nothing written.  Marching-cubes parity stays UNPINNED.' % e)
        return 2
    data = {'skimage_version': np.array(skimage.__version__)}
    cubes = cube_volumes()
    data['cube_volumes'] = cubes
    soups = [soup_of(measure, v) for v in cubes]
    data['cube_counts'] = np.array([len(s) for s in soups], np.int32)
    data['cube_soup'] = np.concatenate(soups) if soups else np.zeros((0, 3), np.float32)
    for k, v in enumerate(noise_volumes()):
        data['noise_volume_%d' % k] = v
        data['noise_soup_%d' % k] = soup_of(measure, v)
    for name, v in config_volumes().items():
        data['config_volume_' + name] = v
        data['config_soup_' + name] = soup_of(measure, v)
    path = os.path.join(HERE, 'mc_skimage.npz')
    np.savez_compressed(path, **data)
    print('wrote %s (skimage %s): %d cube soups, 8 noise volumes, %d config batches'
          % (path, skimage.__version__, len(soups), len([k for k in data if k.startswith('config_soup_')])))
    return 0


if __name__ == '__main__':
    sys.exit(main())

"""Every kernel body once under AddressSanitizer + UndefinedBehaviorSanitizer (TEST INFRASTRUCTURE; started by
tests/test_hostsim_cpu.py with libasan preloaded).  The host execution's answer to compute-sanitizer memcheck:
`__shared__` arrays are statics with redzones, "device" buffers are heap blocks with redzones, so an index one past
a tile, a mask row, a work list or a staged table is reported instead of silently reading a neighbour."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]

import models                                    # noqa: E402
import sdf_b200 as S                             # noqa: E402
import util_hostsim as U                         # noqa: E402
from sdf_b200 import core                        # noqa: E402
from sdf_b200.ir import lower                    # noqa: E402

total = 0
CASES = (('sphere', 33, 32, ('exact', 'f64')), ('csg', 65, 32, ('exact', 'f32')), ('gearlike', 70, 40, ('mixed',)))
for name, n, bs, precisions in CASES:
    prog = lower(models.CONFIGS[name](S))
    hs = U.HostSim(prog, sanitize=True)
    _, L = models.CONFIG_GRIDS[name]
    X = np.arange(-L, L, 2 * L / n)
    tau = core.mixed_tau(prog, X, X, X)
    for prec in precisions:
        for sparse in (True, False) if (name, prec) == ('csg', 'exact') else (True,):
            tris, st = hs.generate(X, X, X, bs, sparse, prec, tau)
            total += len(tris)
    hs.generate(X, X, X, bs, True, 'exact', tau, cap=17)              # a result buffer that is too small
    hs.skip_flags(X, X, X, bs)
    P = models.random_points(5000, seed=1).astype(np.float32)
    hs.eval_f32(P[1:4100])                                            # unaligned, ragged: the plain-load path
    hs.eval_f32(P, fast=True, grid=1)                                 # every stage of the bulk-copy ring reused
    hs.eval_f64(P.astype(np.float64))
rng = np.random.default_rng(1)
ax = np.arange(33.0)
total += len(hs.mc_volume(rng.normal(size=(33, 33, 33)).astype(np.float32), ax, ax, ax))
total += len(hs.mc_volume(rng.normal(size=(5, 33, 2)).astype(np.float32), np.arange(5.0), ax, np.arange(2.0)))
total += len(hs.stl_records(rng.normal(size=(300, 9)).astype(np.float32)))
print('SANITIZED RUN COMPLETE: %d triangles / records' % total)

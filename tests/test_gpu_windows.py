"""sdfb200_generate_window on ONE GPU: the union of windows of the Z-major batch order -- cut at arbitrary batches, inside
a layer -- is the single-call soup, and the batch statistics add up (what N ranks compute in parallel, run here one
after the other).  Needs a B200: -m gpu."""
import numpy as np
import pytest

import models
from util_cmp import compare_soups

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name,n', [('csg', 128), ('gearlike', 192)])
def test_union_of_windows_equals_the_whole_lattice(name, n):
    import sdf_b200 as S
    from sdf_b200 import core, multi
    f = models.CONFIGS[name](S)
    _, L = models.CONFIG_GRIDS[name]
    step, bounds = 2.0 * L / n, ((-L,) * 3, (L,) * 3)
    whole, info = S.generate(f, step=step, bounds=bounds, verbose=False, return_info=True)
    _, _, (X, Y, Z) = core.make_lattice(f, step, bounds)
    X, Y, Z = (np.ascontiguousarray(a, np.float64) for a in (X, Y, Z))
    h, _ = core.program_handle(f)
    tau = core.mixed_tau(h.program, X, Y, Z)
    nb = [(len(a) + 31) // 32 for a in (X, Y, Z)]
    total = nb[0] * nb[1] * nb[2]
    flags = h.skip_flags(X, Y, Z, 32)
    wins = multi.partition_windows(multi.zmajor_active(flags).astype(np.float64), 3)
    assert wins[0][0] == 0 and wins[-1][1] == total and any(w[1] % (nb[0] * nb[1]) for w in wins[:-1])   # a cut inside a layer
    parts, stats = [], np.zeros(3, np.int64)
    for w in wins:
        need, _, _ = h.generate_window(X, Y, Z, 32, True, 'exact', tau, w)
        out = np.empty((need + 8, 9), np.float32)
        n_tri, st, _ = h.generate_window(X, Y, Z, 32, True, 'exact', tau, w, tri_out=out, capacity=len(out))
        assert n_tri == need
        parts.append(out[:n_tri])
        stats += st[:3]
    assert tuple(int(v) for v in stats) == (info['skipped'], info['empty'], info['nonempty'])
    rep = compare_soups(np.concatenate(parts).reshape(-1, 3), whole)
    assert rep['count_equal'] and rep['exact_triangles'] == rep['n_ref'], rep

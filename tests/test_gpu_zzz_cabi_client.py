"""The C-ABI from a plain C99 program (tests/cabi_client.c: hand-built program image, axes, count pass, page-locked
buffer, capacity-overflow check, the real call) against the Python path on the same lattice.  Needs a B200: -m gpu."""
import re

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('n,batch', [(64, 32), (80, 32), (64, 64)])
def test_c_client_result_equals_the_python_path(tmp_path, n, batch):
    import subprocess
    import sdf_b200 as S
    from util_cabi import build_c_client
    exe = build_c_client(tmp_path)
    out = subprocess.run([exe, str(n), str(batch)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith('OK '), out.stdout + out.stderr
    got = {k: float(v) for k, v in re.findall(r'(\w+)=([-+0-9.eE]+)', out.stdout)}
    pts, info = S.generate(S.sphere(1), step=2.5 / n, bounds=((-1.25,) * 3, (1.25,) * 3), batch_size=batch,
                           precision='f64', verbose=False, return_info=True)
    assert info['dims'] == (n, n, n)
    assert (got['n_tri'], got['skipped'], got['empty'], got['nonempty']) == \
        (len(pts) // 3, info['skipped'], info['empty'], info['nonempty'])
    p = pts.astype(np.float64)
    r2 = (p * p).sum(axis=1)
    assert got['n_tri'] > 1000
    assert abs(got['sum'] - p.sum()) <= 1e-9 * max(1.0, np.abs(p).sum())
    assert abs(got['sumsq'] - r2.sum()) <= 1e-9 * r2.sum()
    assert abs(got['r2min'] - r2.min()) <= 1e-8 and abs(got['r2max'] - r2.max()) <= 1e-8
    assert 0.95 < got['r2min'] <= got['r2max'] < 1.05               # vertices of the unit sphere

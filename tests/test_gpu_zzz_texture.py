"""examples/image.py-shaped trees through generate(): text() / image() are host nodes (sdf_b200/texture.py), the
library operands around them evaluate on the GPU, every batch volume is meshed on the device (sdfb200_mc_volume).
Checked against the oracle, which evaluates the same tree in numpy (pinned on reference vectors for these trees in
tests/test_texture_cpu.py) and runs its own marching cubes.  Needs a B200: -m gpu."""
import os

import numpy as np
import pytest

from util_cmp import compare_soups

pytestmark = pytest.mark.gpu

pytest.importorskip('PIL')
pytest.importorskip('scipy.ndimage')


def _mask():
    return np.load(os.path.join(os.path.dirname(__file__), 'golden', 'texture.npz'))['mask_blobs']


def test_image_relief_on_a_plate_matches_the_oracle(tmp_path):
    import sdf_b200 as S
    from oracle import sdf_ref
    w, h = S.measure_image(_mask())
    f = S.rounded_box((w * 1.1, h * 1.1, 0.1), 0.05) | (S.image(_mask()).extrude(1) & S.slab(z0=0, z1=0.075))
    # lattice planes kept off the plate's faces and the slab's planes; 65 samples along X: a one-plane tail batch
    bounds = ((-0.41, -0.61, -0.11), (0.413, 0.6, 0.13))
    kw = dict(step=0.0127, bounds=bounds)
    got, info = S.generate(f, verbose=False, return_info=True, **kw)
    ref, st = sdf_ref.generate(f, **kw)
    assert (info['skipped'], info['empty'], info['nonempty']) == (st['skipped'], st['empty'], st['nonempty']), (info, st)
    rep = compare_soups(got, ref)
    assert rep['n_ref'] > 5000 and rep['count_equal'] and rep['max_vertex_dist'] <= 1e-5 \
        and rep['exact_triangles'] >= 0.999 * rep['n_ref'], rep
    path = os.path.join(str(tmp_path), 'image.stl')
    f.save(path, verbose=False, **kw)
    assert os.path.getsize(path) == 84 + 50 * (len(got) // 3)

#!/usr/bin/env python
"""Generates tests/golden/texture.npz by importing the REAL reference (/root/reference/sdf, sdf/text.py).

Build container only.  Pillow and scipy are installed here, so ``image()`` / ``text()`` / ``measure_*`` of the
reference run as they are; the vectors pin sdf_b200/texture.py (the closures) and the host composition of
extrude / extrude_to / revolve / 2-D transforms / CSG around them (core._compose_around_closure).  Inputs are
synthetic masks built below (stored in the file), and -- when the image has one -- a TrueType file that ships with
the CUDA toolkit documentation, so the test can run wherever this image runs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402

FONT = '/usr/local/cuda-12.9/compute-sanitizer/docs/_static/css/fonts/fontawesome-webfont.ttf'


def masks():
    """Two grey-level test pictures: blobs with a hole (tall) and soft-edged bars (wide; exercises the dither of
    ``convert('1')`` and the down-scaling branch)."""
    yy, xx = np.mgrid[0:96, 0:64]
    a = ((xx - 30) ** 2 + (yy - 40) ** 2 < 20 ** 2) & ~((xx - 34) ** 2 + (yy - 44) ** 2 < 7 ** 2)
    a |= (abs(xx - 40) < 6) & (abs(yy - 78) < 10)
    blobs = (a * 255).astype(np.uint8)
    yy, xx = np.mgrid[0:80, 0:200]
    bars = (127.5 + 127.5 * np.sin(xx / 9.0) * np.cos(yy / 17.0)).astype(np.uint8)
    return dict(blobs=blobs, bars=bars)


def points(rng, n, dim, span):
    p = rng.uniform(-span, span, (n, dim))
    p[:16] = rng.uniform(-1e-3, 1e-3, (16, dim))
    return p


def main():
    ref = import_reference()
    rng = np.random.default_rng(7)
    out = {}
    m = masks()
    p2 = points(rng, 4000, 2, 0.9)
    p3 = points(rng, 4000, 3, 0.9)
    out['p2'], out['p3'] = p2, p3
    for name, arr in m.items():
        out['mask_' + name] = arr
    cases = {
        'blobs_default': lambda r, a: r.image(a['blobs']),
        'blobs_width': lambda r, a: r.image(a['blobs'], width=1.2),
        'bars_height': lambda r, a: r.image(a['bars'], height=0.5),
        'bars_small': lambda r, a: r.image(a['bars'], width=1.5, pixels=3000),
    }
    for key, make in cases.items():
        out['img2_' + key] = make(ref, m)(p2).reshape(-1)
    trees = {
        'extrude': lambda r, a: r.image(a['blobs']).extrude(0.3),
        'example': lambda r, a: r.rounded_box((0.8, 1.2, 0.1), 0.05) | (r.image(a['blobs']).extrude(1) & r.slab(z0=0, z1=0.075)),
        'carved': lambda r, a: r.box((1.2, 0.7, 0.3)) - r.image(a['bars'], width=1.0).extrude(1).translate((0, 0, 0.2)),
        'moved2d': lambda r, a: r.image(a['blobs']).rotate(0.4).translate((0.1, -0.2)).scale(0.8).extrude(0.25),
        'revolve': lambda r, a: r.image(a['blobs'], height=0.6).translate((0.5, 0)).revolve(),
        'extrude_to': lambda r, a: r.image(a['blobs']).extrude_to(r.d2.circle(0.3), 0.6, r.ease.in_out_quad),
        'union2d': lambda r, a: (r.image(a['blobs']) | r.d2.circle(0.2).translate((0.3, 0.3)).k(0.1)).extrude(0.2).orient(r.X),
    }
    for key, make in trees.items():
        out['tree_' + key] = make(ref, m)(p3).reshape(-1)
    out['measure_image'] = np.array([ref.measure_image(m['bars']), ref.measure_image(m['bars'], width=3),
                                     ref.measure_image(m['blobs'], height=2)], np.float64).reshape(-1)
    if os.path.exists(FONT):
        out['text2'] = ref.text(FONT, 'Hi').__call__(p2).reshape(-1)
        out['text2_small'] = ref.text(FONT, 'sdf', width=1.5, pixels=20000, points=128)(p2).reshape(-1)
        out['text_tree'] = (ref.rounded_box((1.4, 1.0, 0.2), 0.05) - ref.text(FONT, 'Hi', height=0.5).extrude(1))(p3).reshape(-1)
        out['measure_text'] = np.array([ref.measure_text(FONT, 'Hi'), ref.measure_text(FONT, 'sdf', width=2)], np.float64).reshape(-1)
    path = os.path.join(HERE, 'texture.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()

"""Model corpus shared by the parity tests, the golden-vector generator and bench.py.

``build(ns)`` builds every model from the names in ``ns`` -- either the reference
package (``import sdf``, only inside tests/golden/make_golden.py) or ``sdf_b200`` --
so one definition drives both sides.  The first five are BASELINE.json's configs
(reference examples/example.py:3-6, gearlike.py:3-5, blobby.py:3-7, knurling.py:4-18).
"""
import math

import numpy as np


def cfg_sphere(ns):
    return ns.sphere(1)


def cfg_csg(ns):                     # examples/example.py
    f = ns.sphere(1) & ns.box(1.5)
    c = ns.cylinder(0.5)
    f -= c.orient(ns.X) | c.orient(ns.Y) | c.orient(ns.Z)
    return f


def cfg_gearlike(ns):                # examples/gearlike.py
    f = ns.sphere(2) & ns.slab(z0=-0.5, z1=0.5).k(0.1)
    f -= ns.cylinder(1).k(0.1)
    f -= ns.cylinder(0.25).circular_array(16, 2).k(0.1)
    return f


def cfg_blobby(ns):                  # examples/blobby.py
    s = ns.sphere(0.75)
    s = s.translate(ns.Z * -3) | s.translate(ns.Z * 3)
    s = s.union(ns.capsule(ns.Z * -3, ns.Z * 3, 0.5), k=1)
    return ns.sphere(1.5).union(s.orient(ns.X), s.orient(ns.Y), s.orient(ns.Z), k=1)


def cfg_knurling(ns):                # examples/knurling.py
    pi = math.pi
    f = ns.rounded_cylinder(1, 0.1, 5)
    x = ns.box((1, 1, 4)).rotate(pi / 4)
    x = x.circular_array(24, 1.6)
    x = x.twist(0.75) | x.twist(-0.75)
    f -= x.k(0.1)
    f -= ns.cylinder(0.5).k(0.1)
    c = ns.cylinder(0.25).orient(ns.X)
    f -= c.translate(ns.Z * -2.5).k(0.1)
    f -= c.translate(ns.Z * 2.5).k(0.1)
    return f


CONFIGS = {
    'sphere': cfg_sphere, 'csg': cfg_csg, 'gearlike': cfg_gearlike,
    'blobby': cfg_blobby, 'knurling': cfg_knurling,
}

# synthetic cubic lattices of BASELINE.md section 3: bounds = +-L, step = 2L/N
CONFIG_GRIDS = {'sphere': (64, 1.25), 'csg': (256, 1.0), 'gearlike': (512, 2.5),
                'blobby': (768, 4.5), 'knurling': (1024, 2.75)}


def node_corpus(ns):
    """One model per node of the 3-D library (d3.py:92-433, dn.py:7-114) that is lowered."""
    X, Y, Z = ns.X, ns.Y, ns.Z
    m = {}
    m['sphere_c'] = ns.sphere(0.7, (0.1, -0.2, 0.3))
    m['plane'] = ns.plane((1, 2, 3), (0.1, 0.2, 0.3))
    m['slab'] = ns.slab(x0=-0.5, x1=0.75, y0=-0.3, z1=0.4)
    m['slab_k'] = ns.slab(z0=-0.5, z1=0.5, x0=0, k=0.2)
    m['box'] = ns.box((1, 2, 3), (0.2, 0.1, -0.1))
    m['box_ab'] = ns.box(a=(-1, -0.5, 0), b=(0.5, 1, 2))
    m['rounded_box'] = ns.rounded_box((1, 2, 3), 0.25)
    m['wireframe_box'] = ns.wireframe_box((1, 2, 3), 0.1)
    m['torus'] = ns.torus(1, 0.25)
    m['capsule'] = ns.capsule(-Z, Z + X * 0.5, 0.5)
    m['cylinder'] = ns.cylinder(0.5)
    m['capped_cylinder'] = ns.capped_cylinder(-Z, Z + Y * 0.25, 0.5)
    m['rounded_cylinder'] = ns.rounded_cylinder(0.5, 0.1, 2)
    m['capped_cone'] = ns.capped_cone(-Z, Z, 1, 0.5)
    m['rounded_cone'] = ns.rounded_cone(0.75, 0.25, 2)
    m['ellipsoid'] = ns.ellipsoid((1, 2, 3))
    m['pyramid'] = ns.pyramid(1)
    m['tetrahedron'] = ns.tetrahedron(1)
    m['octahedron'] = ns.octahedron(1)
    m['dodecahedron'] = ns.dodecahedron(1)
    m['icosahedron'] = ns.icosahedron(1)
    m['translate'] = ns.sphere(0.5).translate((0.3, -0.4, 0.5))
    m['scale'] = ns.box(1).scale(1.5)
    m['scale_xyz'] = ns.sphere(1).scale((1, 2, 3))
    m['rotate'] = ns.capped_cylinder(-Z, Z, 0.5).rotate(math.pi / 4, X)
    m['rotate_axis'] = ns.box((1, 2, 3)).rotate(0.7, (1, 1, 0.5))
    m['orient_x'] = ns.capped_cylinder(-Z, Z, 0.5).orient(X)
    m['orient_negz'] = ns.capped_cone(-Z, Z, 1, 0.5).orient(-Z)
    m['orient_up'] = ns.box((1, 2, 3)).orient(Z)
    m['circular_array'] = ns.capped_cylinder(-Z, Z, 0.25).circular_array(8, 1.5)
    m['elongate'] = ns.sphere(0.5).elongate((0.25, 0.5, 0.75))
    m['twist'] = ns.box((1, 1, 3)).twist(math.pi / 2)
    m['bend'] = ns.box((2, 0.5, 0.5)).bend(0.5)
    m['union'] = ns.sphere(1) | ns.box(1.5)
    m['union3'] = ns.union(ns.sphere(0.6), ns.box(0.9).translate((0.5, 0, 0)), ns.cylinder(0.2))
    m['difference'] = ns.sphere(1) - ns.box((0.5, 0.5, 3))
    m['intersection'] = ns.sphere(1) & ns.box(1.5)
    m['union_k'] = ns.sphere(1) | ns.box(1.5).translate((0.8, 0, 0)).k(0.25)
    m['difference_k'] = ns.difference(ns.sphere(1), ns.box((0.5, 0.5, 3)), k=0.25)
    m['intersection_k'] = ns.sphere(1) & ns.box(1.5).k(0.25)
    m['blend'] = ns.sphere(1).blend(ns.box(1.5), k=0.3)
    m['negate'] = ns.sphere(1).negate()
    m['dilate'] = ns.box(1).dilate(0.25)
    m['erode'] = ns.box(2).erode(0.25)
    m['shell'] = ns.sphere(1).shell(0.1) & ns.plane()
    m['repeat'] = ns.sphere(0.4).repeat(1.0)
    m['repeat_count'] = ns.sphere(0.4).repeat((1.0, 1.5, 0), count=(2, 1, 0), padding=1)
    m['nested'] = (ns.sphere(1).translate((0.2, 0, 0)) | ns.box(0.8).rotate(0.3, Y).k(0.1)).shell(0.05).twist(0.5) \
        - ns.cylinder(0.3).orient(X).translate((0, 0.1, 0.2))
    return m


def build(ns):
    out = {('cfg_' + k): f(ns) for k, f in CONFIGS.items()}
    out.update(node_corpus(ns))
    return out


def lattice_for(name):
    n, L = CONFIG_GRIDS[name]
    return ((-L, -L, -L), (L, L, L)), 2.0 * L / n


def random_points(n, seed=0, L=3.0):
    rng = np.random.default_rng(seed)
    p = rng.uniform(-L, L, (n, 3))
    # exact lattice-like coordinates and zeros exercise the branch points (atan2(0,0), |x| = 0, ...)
    p[:8] = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (-1, 0, 0), (0.5, 0.5, 0.5), (2, 0, 0), (0, -2, 0)]
    return p

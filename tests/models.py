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


# Triangle counts of the ORACLE (numpy evaluation + C Lewiner restatement, tests/test_oracle_cpu.py) on the
# reference-natural grids of configs 3-5 (samples = N^3, reference bounds); the GPU path must reproduce them.
NATURAL_TRIANGLES = {'gearlike': 2550892, 'blobby': 2278232, 'knurling': 18721558}


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


EASE_NAMES = ['linear', 'in_quad', 'out_quad', 'in_out_quad', 'in_cubic', 'out_cubic', 'in_out_cubic',
              'in_quart', 'out_quart', 'in_out_quart', 'in_quint', 'out_quint', 'in_out_quint',
              'in_sine', 'out_sine', 'in_out_sine', 'in_expo', 'out_expo', 'in_out_expo',
              'in_circ', 'out_circ', 'in_out_circ', 'in_elastic', 'out_elastic', 'in_out_elastic',
              'in_back', 'out_back', 'in_out_back', 'in_bounce', 'out_bounce', 'in_out_bounce',
              'in_square', 'out_square', 'in_out_square']


def example_pawn(ns):                # examples/pawn.py
    Z, ease = ns.Z, ns.ease

    def section(z0, z1, d0, d1, e=ease.linear):
        f = ns.cylinder(d0 / 2).transition_linear(ns.cylinder(d1 / 2), Z * z0, Z * z1, e)
        return f & ns.slab(z0=z0, z1=z1)
    f = section(0, 0.2, 1, 1.25)
    f |= section(0.2, 0.3, 1.25, 1).k(0.05)
    f |= ns.rounded_cylinder(0.6, 0.1, 0.2).translate(Z * 0.4).k(0.05)
    f |= section(0.5, 1.75, 1, 0.25, ease.out_quad).k(0.01)
    f |= section(1.75, 1.85, 0.25, 0.5).k(0.01)
    f |= section(1.85, 1.90, 0.5, 0.25).k(0.05)
    f |= ns.sphere(0.3).translate(Z * 2.15).k(0.05)
    return f


def example_weave(ns):               # examples/weave.py
    X, Z, ease = ns.X, ns.Z, ns.ease
    f = ns.rounded_box([3.2, 1, 0.25], 0.1).translate((1.5, 0, 0.0625))
    f = f.bend_linear(X * 0.75, X * 2.25, Z * -0.1875, ease.in_out_quad)
    f = f.circular_array(3, 0)
    f = f.repeat((2.7, 5.4, 0), padding=1)
    f |= f.translate((2.7 / 2, 2.7, 0))
    f &= ns.cylinder(10)
    f |= (ns.cylinder(12) - ns.cylinder(10)) & ns.slab(z0=-0.5, z1=0.5).k(0.25)
    return f


def widened_corpus(ns):
    """the 2-D library, extrude / extrude_to / revolve, slice, the
    eased warps and transitions, every easing function (d2.py, d3.py:435-520, ease.py)."""
    X, Y, Z, ease = ns.X, ns.Y, ns.Z, ns.ease
    m = {}
    m['x_hexagon'] = ns.hexagon(1).extrude(0.5)
    m['x_rectangle'] = ns.rectangle((1, 2), (0.1, -0.2)).extrude(1)
    m['x_rectangle_ab'] = ns.rectangle(a=(-1, -0.5), b=(0.5, 1)).extrude(0.4)
    m['x_rounded_rectangle'] = ns.rounded_rectangle(np.array((1.0, 2.0)), (0.1, 0.2, 0.3, 0.05)).extrude(0.5)
    m['x_triangle'] = ns.equilateral_triangle().extrude(0.3)
    m['x_rounded_x'] = ns.rounded_x(1, 0.1).extrude(0.2)
    m['x_polygon'] = ns.polygon([(0, 0), (1, 0), (1.2, 0.8), (0.5, 1.5), (-0.3, 0.7)]).extrude(0.4)
    m['x_vesica'] = ns.vesica(1, 0.3).extrude(0.2)
    m['x_line_slab'] = ns.d2.slab(x0=-0.5, x1=0.5, y0=-0.2).extrude(0.5)
    m['revolve'] = ns.circle(0.4).revolve(1.2)
    m['revolve_hex'] = ns.hexagon(0.3).translate((0.1, 0.2)).revolve(1.0)
    m['extrude_to'] = ns.rectangle(1).extrude_to(ns.circle(0.7), 1.5, ease.in_out_quad)
    m['ops2d'] = (ns.circle(0.5).translate((0.3, 0.1)) | ns.rectangle((0.6, 0.4)).rotate(0.5).k(0.1)).scale(1.3).extrude(0.5)
    m['scale2d_xy'] = ns.circle(0.5).scale((1.0, 2.0)).extrude(0.3)
    m['circular_array2d'] = ns.rectangle((0.2, 0.8)).translate((0.8, 0)).circular_array(5).extrude(0.3)
    m['elongate2d'] = ns.circle(0.3).elongate((0.4, 0.2)).extrude(0.3)
    m['repeat2d'] = ns.circle(0.2).repeat((0.6, 0.6), padding=1).extrude(0.2)
    m['csg2d'] = ((ns.circle(1) & ns.rectangle(1.5)) - ns.circle(0.4).k(0.1)).shell(0.1).dilate(0.02).extrude(0.4)
    m['slice'] = (ns.sphere(1) | ns.box(1.2).translate((0.5, 0, 0.3))).slice().extrude(0.3)
    m['bend_linear'] = ns.rounded_box([3.2, 1, 0.25], 0.1).translate((1.5, 0, 0.0625)).bend_linear(
        X * 0.75, X * 2.25, Z * -0.1875, ease.in_out_quad)
    m['bend_radial'] = ns.box((2, 2, 0.2)).bend_radial(0.2, 1.0, 0.5, ease.in_out_sine)
    m['transition_radial'] = ns.box(1).transition_radial(ns.sphere(0.8), 0.2, 0.9, ease.out_cubic)
    m['transition_default'] = ns.sphere(0.8).transition_linear(ns.box(1.2))
    m['wrap_around'] = ns.box((4, 0.3, 0.5)).translate((0, 1.0, 0)).wrap_around(-2, 2)
    m['wrap_around_r'] = ns.capsule(-X * 2, X * 2, 0.2).wrap_around(-2, 2, r=0.8, e=ease.in_out_cubic)
    m['pawn'] = example_pawn(ns)
    m['weave'] = example_weave(ns)
    for name in EASE_NAMES:
        m['ease_' + name] = ns.sphere(0.8).transition_linear(ns.box(1.2), -Z * 0.7, Z * 0.9, getattr(ease, name))
    return m


def build(ns):
    out = {('cfg_' + k): f(ns) for k, f in CONFIGS.items()}
    out.update(node_corpus(ns))
    out.update(widened_corpus(ns))
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


def random_tree(S, rng, depth):
    """A seeded random composition of primitives, transforms and folds (S = our package or the reference)."""
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
    a = random_tree(S, rng, depth - 1)
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
    b = random_tree(S, rng, depth - 1)
    kk = None if rng.random() < 0.5 else u(0.05, 0.4)
    if k == 11: return a.union(b, random_tree(S, rng, depth - 1), k=kk)
    if k == 12: return a.difference(b, k=kk)
    if k == 13: return a.intersection(b, k=kk)
    if k == 14: return a.blend(b, k=u(0.1, 0.9))
    return a.transition_linear(b, v3(-1, 0), v3(0, 1))


def random_trees(S, count=120, seed=424242):
    """The fixed list of random trees behind tests/golden/random_trees.npz."""
    rng = np.random.default_rng(seed)
    return [random_tree(S, rng, int(rng.integers(1, 5))) for _ in range(count)]

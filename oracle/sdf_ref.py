"""ORACLE (test infrastructure, not product code).

CPU restatement, in numpy float64, of the reference's mesh-generation hot path:

  * node formulas            /root/reference/sdf/d3.py:92-502, sdf/dn.py:7-114
  * grid / batching / skip   sdf/core.py:20-60, 84-150
  * bounds estimation        sdf/core.py:62-82
  * marching cubes           sdf/core.py:16-18 -> oracle/mc_lewiner.c (C restatement of
                             skimage's Lewiner MC; PARITY UNPINNED, see that file)

It walks the product's IR tree (sdf_b200.ir.Node) so that one model definition
drives both the CUDA path and this checker.  The numpy expressions follow the
reference's operation order so that, in float64, results are bit-identical to
the reference closures wherever numpy itself is deterministic (pinned by
tests/golden/*.npz, which were produced by importing the real reference with a
stub ``skimage`` module: tests/golden/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
import ctypes
import itertools
import os
import subprocess

import numpy as np

from sdf_b200.ir import node_of

_HERE = os.path.dirname(os.path.abspath(__file__))

_min = np.minimum
_max = np.maximum


def _length(a):
    return np.linalg.norm(a, axis=1)


def _dot(a, b):
    return np.sum(a * b, axis=1)


def _vec(*arrs):
    return np.stack(arrs, axis=-1)


# --------------------------------------------------------------------------
# node evaluation.  Every function takes the constants (float64 list) and the
# (N,3) point block and returns what the reference closure returns.

def _p_sphere(c, p):                       # d3.py:92-96
    return _length(p - np.array(c[0:3])) - c[3]


def _p_plane(c, p):                        # d3.py:98-103
    return np.dot(np.array(c[3:6]) - p, np.array(c[0:3]))


def _p_box(c, p):                          # d3.py:131-133 (c[3:6] is size / 2)
    q = np.abs(p - np.array(c[0:3])) - np.array(c[3:6])
    return _length(_max(q, 0)) + _min(np.amax(q, axis=1), 0)


def _p_rounded_box(c, p):                  # d3.py:139-141
    q = np.abs(p) - np.array(c[0:3]) + c[3]
    return _length(_max(q, 0)) + _min(np.amax(q, axis=1), 0) - c[3]


def _p_wireframe_box(c, p):                # d3.py:147-154
    half, t = np.array(c[0:3]), c[3]

    def g(a, b, cc):
        return _length(_max(_vec(a, b, cc), 0)) + _min(_max(a, _max(b, cc)), 0)
    p = np.abs(p) - half - t / 2
    q = np.abs(p + t / 2) - t / 2
    px, py, pz = p[:, 0], p[:, 1], p[:, 2]
    qx, qy, qz = q[:, 0], q[:, 1], q[:, 2]
    return _min(_min(g(px, qy, qz), g(qx, py, qz)), g(qx, qy, pz))


def _p_torus(c, p):                        # d3.py:159-164
    a = _length(p[:, [0, 1]]) - c[0]
    return _length(_vec(a, p[:, 2])) - c[1]


def _p_capsule(c, p):                      # d3.py:171-175
    a, b = np.array(c[0:3]), np.array(c[3:6])
    pa = p - a
    ba = b - a
    h = np.clip(np.dot(pa, ba) / np.dot(ba, ba), 0, 1).reshape((-1, 1))
    return _length(pa - np.multiply(ba, h)) - c[6]


def _p_cylinder(c, p):                     # d3.py:180-181
    return _length(p[:, [0, 1]]) - c[0]


def _p_capped_cylinder(c, p):              # d3.py:188-203
    a, b, radius = np.array(c[0:3]), np.array(c[3:6]), c[6]
    ba = b - a
    pa = p - a
    baba = np.dot(ba, ba)
    paba = np.dot(pa, ba).reshape((-1, 1))
    x = _length(pa * baba - ba * paba) - radius * baba
    y = np.abs(paba - baba * 0.5) - baba * 0.5
    x = x.reshape((-1, 1))
    y = y.reshape((-1, 1))
    x2 = x * x
    y2 = y * y * baba
    d = np.where(_max(x, y) < 0, -_min(x2, y2), np.where(x > 0, x2, 0) + np.where(y > 0, y2, 0))
    return np.sign(d) * np.sqrt(np.abs(d)) / baba


def _p_rounded_cylinder(c, p):             # d3.py:208-214
    ra, rb, h = c
    d = _vec(_length(p[:, [0, 1]]) - ra + rb, np.abs(p[:, 2]) - h / 2 + rb)
    return _min(_max(d[:, 0], d[:, 1]), 0) + _length(_max(d, 0)) - rb


def _p_capped_cone(c, p):                  # d3.py:221-236
    a, b, ra, rb = np.array(c[0:3]), np.array(c[3:6]), c[6], c[7]
    rba = rb - ra
    baba = np.dot(b - a, b - a)
    papa = _dot(p - a, p - a)
    paba = np.dot(p - a, b - a) / baba
    x = np.sqrt(papa - paba * paba * baba)
    cax = _max(0, x - np.where(paba < 0.5, ra, rb))
    cay = np.abs(paba - 0.5) - 0.5
    k = rba * rba + baba
    f = np.clip((rba * (x - ra) + paba * baba) / k, 0, 1)
    cbx = x - ra - f * rba
    cby = paba - f
    s = np.where(np.logical_and(cbx < 0, cay < 0), -1, 1)
    return s * np.sqrt(_min(cax * cax + cay * cay * baba, cbx * cbx + cby * cby * baba))


def _p_rounded_cone(c, p):                 # d3.py:241-249
    r1, r2, h = c
    q = _vec(_length(p[:, [0, 1]]), p[:, 2])
    b = (r1 - r2) / h
    a = np.sqrt(1 - b * b)
    k = np.dot(q, _vec(-b, a))
    c1 = _length(q) - r1
    c2 = _length(q - _vec(0, h)) - r2
    c3 = np.dot(q, _vec(a, b)) - r1
    return np.where(k < 0, c1, np.where(k > a * h, c2, c3))


def _p_ellipsoid(c, p):                    # d3.py:255-258
    size = np.array(c[0:3])
    k0 = _length(p / size)
    k1 = _length(p / (size * size))
    return k0 * (k0 - 1) / k1


def _p_pyramid(c, p):                      # d3.py:263-281
    h = c[0]
    a = np.abs(p[:, [0, 1]]) - 0.5
    w = a[:, 1] > a[:, 0]
    a[w] = a[:, [1, 0]][w]
    px = a[:, 0]
    py = p[:, 2]
    pz = a[:, 1]
    m2 = h * h + 0.25
    qx = pz
    qy = h * py - 0.5 * px
    qz = h * px + 0.5 * py
    s = _max(-qx, 0)
    t = np.clip((qy - 0.5 * pz) / (m2 + 0.25), 0, 1)
    a = m2 * (qx + s) ** 2 + qy * qy
    b = m2 * (qx + 0.5 * t) ** 2 + (qy - m2 * t) ** 2
    d2 = np.where(_min(qy, -qx * m2 - qy * 0.5) > 0, 0, _min(a, b))
    return np.sqrt((d2 + qz * qz) / m2) * np.sign(_max(qz, -py))


def _p_tetrahedron(c, p):                  # d3.py:288-292
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    return (_max(np.abs(x + y) - z, np.abs(x - y) + z) - c[0]) / np.sqrt(3)


def _p_octahedron(c, p):                   # d3.py:297-298 (c[1] = tan(radians(30)))
    return (np.sum(np.abs(p), axis=1) - c[0]) * c[1]


def _p_dodecahedron(c, p):                 # d3.py:303-310
    r, x, y, z = c
    p = np.abs(p / r)
    a = np.dot(p, (x, y, z))
    b = np.dot(p, (z, x, y))
    cc = np.dot(p, (y, z, x))
    return (_max(_max(a, b), cc) - x) * r


def _p_icosahedron(c, p):                  # d3.py:318-324 (c[0] already scaled by 0.85065...)
    r, x, y, z, w = c
    p = np.abs(p / r)
    a = np.dot(p, (x, y, z))
    b = np.dot(p, (z, x, y))
    cc = np.dot(p, (y, z, x))
    d = np.dot(p, (w, w, w)) - x
    return _max(_max(_max(a, b), cc) - x, d) * r


_PRIMS = {
    'SPHERE': _p_sphere, 'PLANE': _p_plane, 'BOX': _p_box, 'ROUNDED_BOX': _p_rounded_box,
    'WIREFRAME_BOX': _p_wireframe_box, 'TORUS': _p_torus, 'CAPSULE': _p_capsule,
    'CYLINDER': _p_cylinder, 'CAPPED_CYLINDER': _p_capped_cylinder,
    'ROUNDED_CYLINDER': _p_rounded_cylinder, 'CAPPED_CONE': _p_capped_cone,
    'ROUNDED_CONE': _p_rounded_cone, 'ELLIPSOID': _p_ellipsoid, 'PYRAMID': _p_pyramid,
    'TETRAHEDRON': _p_tetrahedron, 'OCTAHEDRON': _p_octahedron,
    'DODECAHEDRON': _p_dodecahedron, 'ICOSAHEDRON': _p_icosahedron,
}


# ---- 2-D primitives (p is (N,2))                                        sdf/d2.py

def _q_circle(c, p):                        # d2.py:78-79
    return _length(p - np.array(c[0:2])) - c[2]


def _q_line(c, p):                          # d2.py:85-86
    return np.dot(np.array(c[2:4]) - p, np.array(c[0:2]))


def _q_rectangle(c, p):                     # d2.py:111-113 (c[2:4] is size / 2)
    q = np.abs(p - np.array(c[0:2])) - np.array(c[2:4])
    return _length(_max(q, 0)) + _min(np.amax(q, axis=1), 0)


def _q_rounded_rectangle(c, p):             # d2.py:122-134 (c[0:2] is size / 2)
    half, (r0, r1, r2, r3) = np.array(c[0:2]), c[2:6]
    x = p[:, 0]
    y = p[:, 1]
    r = np.zeros(len(p)).reshape((-1, 1))
    r[np.logical_and(x > 0, y > 0)] = r0
    r[np.logical_and(x > 0, y <= 0)] = r1
    r[np.logical_and(x <= 0, y <= 0)] = r2
    r[np.logical_and(x <= 0, y > 0)] = r3
    q = np.abs(p) - half + r
    return (_min(_max(q[:, 0], q[:, 1]), 0).reshape((-1, 1)) + _length(_max(q, 0)).reshape((-1, 1)) - r)


def _q_equilateral_triangle(c, p):          # d2.py:139-152
    k = 3 ** 0.5
    p = _vec(np.abs(p[:, 0]) - 1, p[:, 1] + 1 / k)
    w = p[:, 0] + k * p[:, 1] > 0
    q = _vec(p[:, 0] - k * p[:, 1], -k * p[:, 0] - p[:, 1]) / 2
    p = np.where(w.reshape((-1, 1)), q, p)
    p = _vec(p[:, 0] - np.clip(p[:, 0], -2, 0), p[:, 1])
    return -_length(p) * np.sign(p[:, 1])


def _q_hexagon(c, p):                       # d2.py:157-166 (c[0] already scaled by sqrt(3)/2)
    r = c[0]
    k = np.array((3 ** 0.5 / -2, 0.5, np.tan(np.pi / 6)))
    p = np.abs(p)
    p -= 2 * k[:2] * _min(_dot(k[:2], p), 0).reshape((-1, 1))
    p -= _vec(np.clip(p[:, 0], -k[2] * r, k[2] * r), np.zeros(len(p)) + r)
    return _length(p) * np.sign(p[:, 1])


def _q_rounded_x(c, p):                     # d2.py:171-174
    w, r = c
    p = np.abs(p)
    q = (_min(p[:, 0] + p[:, 1], w) * 0.5).reshape((-1, 1))
    return _length(p - q) - r


def _q_polygon(c, p):                       # d2.py:180-197
    points = [np.array(c[2 * i:2 * i + 2]) for i in range(len(c) // 2)]
    n = len(points)
    d = _dot(p - points[0], p - points[0])
    s = np.ones(len(p))
    for i in range(n):
        j = (i + n - 1) % n
        vi = points[i]
        vj = points[j]
        e = vj - vi
        w = p - vi
        b = w - e * np.clip(np.dot(w, e) / np.dot(e, e), 0, 1).reshape((-1, 1))
        d = _min(d, _dot(b, b))
        c1 = p[:, 1] >= vi[1]
        c2 = p[:, 1] < vj[1]
        c3 = e[0] * w[:, 1] > e[1] * w[:, 0]
        cc = _vec(c1, c2, c3)
        s = np.where(np.all(cc, axis=1) | np.all(~cc, axis=1), -s, s)
    return s * np.sqrt(d)


def _q_vesica(c, p):                        # d2.py:202-208
    r, d = c
    p = np.abs(p)
    b = np.sqrt(r * r - d * d)
    return np.where(((p[:, 1] - b) * d > p[:, 0] * b), _length(p - np.array([0, b])),
                    _length(p - np.array([-d, 0])) - r)


_PRIMS2 = {'CIRCLE': _q_circle, 'LINE2': _q_line, 'RECTANGLE': _q_rectangle,
           'ROUNDED_RECTANGLE': _q_rounded_rectangle, 'EQUILATERAL_TRIANGLE': _q_equilateral_triangle,
           'HEXAGON': _q_hexagon, 'ROUNDED_X': _q_rounded_x, 'POLYGON': _q_polygon, 'VESICA': _q_vesica}


# ---- easing functions, by opcode operand                                sdf/ease.py:3-162

def _out_bounce(t):
    a = (121 * t * t) / 16
    b = (363 / 40 * t * t) - (99 / 10 * t) + 17 / 5
    c = (4356 / 361 * t * t) - (35442 / 1805 * t) + 16061 / 1805
    d = (54 / 5 * t * t) - (513 / 25 * t) + 268 / 25
    return np.where(t < 4 / 11, a, np.where(t < 8 / 11, b, np.where(t < 9 / 10, c, d)))


def _in_bounce(t):
    return 1 - _out_bounce(1 - t)


def _w(cond, a, b):
    with np.errstate(all='ignore'):
        return np.where(cond, a, b)


_EASE = [
    lambda t: t,
    lambda t: t * t,
    lambda t: -t * (t - 2),
    lambda t: _w(t < 0.5, 2 * t * t, -0.5 * ((2 * t - 1) * ((2 * t - 1) - 2) - 1)),
    lambda t: t * t * t,
    lambda t: (t - 1) * (t - 1) * (t - 1) + 1,
    lambda t: _w(t * 2 < 1, 0.5 * (t * 2) * (t * 2) * (t * 2), 0.5 * ((t * 2 - 2) * (t * 2 - 2) * (t * 2 - 2) + 2)),
    lambda t: t * t * t * t,
    lambda t: -((t - 1) * (t - 1) * (t - 1) * (t - 1) - 1),
    lambda t: _w(t * 2 < 1, 0.5 * (t * 2) * (t * 2) * (t * 2) * (t * 2),
                 -0.5 * ((t * 2 - 2) * (t * 2 - 2) * (t * 2 - 2) * (t * 2 - 2) - 2)),
    lambda t: t * t * t * t * t,
    lambda t: (t - 1) * (t - 1) * (t - 1) * (t - 1) * (t - 1) + 1,
    lambda t: _w(t * 2 < 1, 0.5 * (t * 2) * (t * 2) * (t * 2) * (t * 2) * (t * 2),
                 0.5 * ((t * 2 - 2) * (t * 2 - 2) * (t * 2 - 2) * (t * 2 - 2) * (t * 2 - 2) + 2)),
    lambda t: -np.cos(t * np.pi / 2) + 1,
    lambda t: np.sin(t * np.pi / 2),
    lambda t: -0.5 * (np.cos(np.pi * t) - 1),
    lambda t: _w(t == 0, np.zeros(len(t)), 2 ** (10 * (t - 1))),
    lambda t: _w(t == 1, np.zeros(len(t)) + 1, 1 - 2 ** (-10 * t)),
    lambda t: _w(t == 0, np.zeros(len(t)), _w(t == 1, np.zeros(len(t)) + 1,
                 _w(t < 0.5, 0.5 * 2 ** (20 * t - 10), 1 - 0.5 * 2 ** (-20 * t + 10)))),
    lambda t: -1 * (np.sqrt(1 - t * t) - 1),
    lambda t: np.sqrt(1 - (t - 1) * (t - 1)),
    lambda t: _w(t * 2 < 1, -0.5 * (_sqrt_q(1 - (t * 2) * (t * 2)) - 1), 0.5 * (_sqrt_q(1 - (t * 2 - 2) * (t * 2 - 2)) + 1)),
    lambda t: -1 * (2 ** (10 * (t - 1)) * np.sin(((t - 1) - 0.5 / 4) * (2 * np.pi) / 0.5)),
    lambda t: 2 ** (-10 * t) * np.sin((t - 0.5 / 4) * (2 * np.pi / 0.5)) + 1,
    lambda t: _w(t * 2 < 1, -0.5 * (2 ** (10 * (t * 2 - 1)) * np.sin(((t * 2 - 1) - 0.5 / 4) * 2 * np.pi / 0.5)),
                 2 ** (-10 * (t * 2 - 1)) * np.sin(((t * 2 - 1) - 0.5 / 4) * 2 * np.pi / 0.5) * 0.5 + 1),
    lambda t: t * t * ((1.70158 + 1) * t - 1.70158),
    lambda t: (t - 1) * (t - 1) * ((1.70158 + 1) * (t - 1) + 1.70158) + 1,
    lambda t: _w(t * 2 < 1, 0.5 * ((t * 2) * (t * 2) * ((1.70158 * 1.525 + 1) * (t * 2) - 1.70158 * 1.525)),
                 0.5 * ((t * 2 - 2) * (t * 2 - 2) * ((1.70158 * 1.525 + 1) * (t * 2 - 2) + 1.70158 * 1.525) + 2)),
    _in_bounce,
    _out_bounce,
    lambda t: _w(t < 0.5, _in_bounce(2 * t) * 0.5, _out_bounce(2 * t - 1) * 0.5 + 0.5),
    lambda t: _w(t < 1, np.zeros(len(t)), np.zeros(len(t)) + 1),
    lambda t: _w(t > 0, np.zeros(len(t)) + 1, np.zeros(len(t))),
    lambda t: _w(t < 0.5, np.zeros(len(t)), np.zeros(len(t)) + 1),
]


def _sqrt_q(v):
    with np.errstate(invalid='ignore'):
        return np.sqrt(v)


def _tparam(tp, p):
    name, c = tp
    if name == 'TPARAM_LINEAR':             # d3.py:442 / 464
        p0, ab = np.array(c[0:3]), np.array(c[3:6])
        return np.clip(np.dot(p - p0, ab) / np.dot(ab, ab), 0, 1)
    if name == 'TPARAM_RADIAL':             # d3.py:453-454 / 476-477
        r = np.hypot(p[:, 0], p[:, 1])
        return np.clip((r - c[0]) / (c[1] - c[0]), 0, 1)
    if name == 'TPARAM_Z':                  # d2.py:273
        return np.clip(p[:, 2] / c[0], -0.5, 0.5) + 0.5
    raise KeyError(name)


def _fold(name, K, d1, d2):
    if name == 'union':                     # dn.py:13-18
        if K is None:
            return _min(d1, d2)
        h = np.clip(0.5 + 0.5 * (d2 - d1) / K, 0, 1)
        m = d2 + (d1 - d2) * h
        return m - K * h * (1 - h)
    if name == 'difference':                # dn.py:28-33
        if K is None:
            return _max(d1, -d2)
        h = np.clip(0.5 - 0.5 * (d2 + d1) / K, 0, 1)
        m = d1 + (-d2 - d1) * h
        return m + K * h * (1 - h)
    if name == 'intersection':              # dn.py:43-48
        if K is None:
            return _max(d1, d2)
        h = np.clip(0.5 - 0.5 * (d2 - d1) / K, 0, 1)
        m = d2 + (d1 - d2) * h
        return m + K * h * (1 - h)
    raise KeyError(name)


def evaluate(obj, p):
    """sdf(p): (N,3) float64 -> (N,1) float64, like SDF3.__call__ (d3.py:24-25)."""
    p = np.asarray(p)
    n = node_of(obj)
    kind, c = n.kind, n.consts
    if kind == 'opaque':
        return obj.f(p).reshape((-1, 1))
    if kind == 'prim' and n.op in _PRIMS2:
        return _PRIMS2[n.op](c, p[:, :2]).reshape((-1, 1))
    if kind == 'polygon':
        return _q_polygon(c, p[:, :2]).reshape((-1, 1))
    if kind == 'prim':
        return _PRIMS[n.op](c, p).reshape((-1, 1))
    child = n.children[0] if n.children else None
    if kind == 'transform':
        if n.op == 'TRANSLATE':             # d3.py:331-332
            return evaluate(child, p - np.array(c))
        if n.op == 'ROTATE':                # d3.py:358-359
            return evaluate(child, np.dot(p, np.array(c).reshape(3, 3)))
        if n.op == 'REVOLVE':               # d2.py:282-285 (2-D child sees (|xy| - offset, z))
            q = _vec(_length(p[:, [0, 1]]) - c[0], p[:, 2])
            return evaluate(child, np.concatenate([q, np.zeros((len(p), 1))], axis=1))
        if n.op in ('TWIST', 'BEND'):       # d3.py:409-418, 423-432
            x, y, z = p[:, 0], p[:, 1], p[:, 2]
            arg = c[0] * (z if n.op == 'TWIST' else x)
            cs, sn = np.cos(arg), np.sin(arg)
            return evaluate(child, _vec(cs * x - sn * y, sn * x + cs * y, z))
        raise KeyError(n.op)
    if kind == 'scale':                     # d3.py:343-344
        s = (c[0], c[1], c[2])
        return evaluate(child, p / s) * c[3]
    if kind == 'elongate':                  # d3.py:398-404
        q = np.abs(p) - np.array(c)
        x = q[:, 0].reshape((-1, 1))
        y = q[:, 1].reshape((-1, 1))
        z = q[:, 2].reshape((-1, 1))
        w = _min(_max(x, _max(y, z)), 0)
        return evaluate(child, _max(q, 0)) + w
    if kind == 'extrude':                   # d2.py:263-265
        d = evaluate(child, p)
        w = _vec(d.reshape(-1), np.abs(p[:, 2]) - c[0] / 2)
        return (_min(_max(w[:, 0], w[:, 1]), 0) + _length(_max(w, 0))).reshape((-1, 1))
    if kind == 'warp':
        if n.op in ('BEND_LINEAR', 'BEND_RADIAL'):      # d3.py:441-444, 449-456
            (tp, _, tc), (_, eid, _), (_, _, v) = n.extra
            t = _EASE[eid](_tparam((tp, tc), p)).reshape((-1, 1))
            if n.op == 'BEND_LINEAR':
                return evaluate(child, p + t * np.array(v))
            x, y, z = p[:, 0], p[:, 1], p[:, 2]
            return evaluate(child, _vec(x, y, z - (-v[2]) * t.reshape(-1)))
        if n.op == 'WRAP_AROUND':                       # d3.py:490-501
            (_, _, (r,)), (_, eid, _), (_, _, (x0, x1)) = n.extra
            p0 = np.array((1, 0, 0)) * x0
            p1 = np.array((1, 0, 0)) * x1
            v = -np.array((0, 1, 0))
            x, y, z = p[:, 0], p[:, 1], p[:, 2]
            d = (np.hypot(x, y) - r).reshape((-1, 1))
            a = np.arctan2(y, x)
            t = _EASE[eid]((a + np.pi) / (2 * np.pi)).reshape((-1, 1))
            q = p0 + (p1 - p0) * t + v * d
            q[:, 2] = z
            return evaluate(child, q)
        raise KeyError(n.op)
    if kind == 'transition':
        tp, eid, mix, tail = n.extra
        d1 = evaluate(n.children[0], p)
        d2 = evaluate(n.children[1], p)
        t = _EASE[eid](_tparam(tp, p)).reshape((-1, 1))
        if mix == 'LERP_POP2':                          # d3.py:466, 479
            return t * d2 + (1 - t) * d1
        d = d1 + (d2 - d1) * t                          # d2.py:274-277 (extrude_to)
        w = _vec(d.reshape(-1), np.abs(p[:, 2]) - tail[0] / 2)
        return (_min(_max(w[:, 0], w[:, 1]), 0) + _length(_max(w, 0))).reshape((-1, 1))
    if kind == 'slice':                     # d3.py:513-519
        q = _vec(p[:, 0], p[:, 1], np.zeros(len(p)))
        A = evaluate(n.children[0], q).reshape(-1)
        B = -evaluate(n.children[1], q).reshape(-1)
        w = A <= 0
        A[w] = B[w]
        return A.reshape((-1, 1))
    if kind == 'circular':                  # d3.py:381-391
        da, offset = c
        shifted = lambda q: evaluate(child, q - np.array((1, 0, 0)) * offset)
        x, y, z = p[:, 0], p[:, 1], p[:, 2]
        d = np.hypot(x, y)
        a = np.arctan2(y, x) % da
        d1 = shifted(_vec(np.cos(a - da) * d, np.sin(a - da) * d, z))
        d2 = shifted(_vec(np.cos(a) * d, np.sin(a) * d, z))
        return _min(d1, d2)
    if kind == 'value':
        d = evaluate(child, p)
        if n.op == 'NEGATE':                # dn.py:63-64
            return -d
        if n.op == 'ADDC':                  # dn.py:68-69, 73-74 (c = -r / +r)
            return d + c[0]
        if n.op == 'SHELL':                 # dn.py:78-79
            return np.abs(d) - c[0] / 2
        raise KeyError(n.op)
    if kind == 'fold':
        d1 = evaluate(n.children[0], p)
        for b in n.children[1:]:
            d2 = evaluate(b, p)
            K = n.extra or getattr(b, '_k', None)
            d1 = _fold(n.op, K, d1, d2)
        return d1
    if kind == 'blend':                     # dn.py:53-59
        d1 = evaluate(n.children[0], p)
        for b in n.children[1:]:
            d2 = evaluate(b, p)
            K = n.extra or getattr(b, '_k', None)
            d1 = K * d2 + (1 - K) * d1
        return d1
    if kind == 'repeat':                    # dn.py:101-113
        spacing, count, has_count, neighbours = n.extra
        dim = p.shape[-1]
        spacing = np.array(spacing[:dim])
        q = np.divide(p, spacing, out=np.zeros_like(p), where=spacing != 0)
        if has_count:
            cnt = np.array(count[:dim])
            index = np.clip(np.round(q), -cnt, cnt)
        else:
            index = np.round(q)
        A = [evaluate(child, p - spacing * (index + np.array(nb[:dim]))) for nb in neighbours]
        a = A[0]
        for b in A[1:]:
            a = _min(a, b)
        return a
    raise KeyError(kind)


# --------------------------------------------------------------------------
# marching cubes through the C oracle

_lib = None


def mc_lib():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, 'libmc_oracle.so')
        src = os.path.join(_HERE, 'mc_lewiner.c')
        if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(['make', '-s', '-C', _HERE])
        _lib = ctypes.CDLL(so)
        _lib.mc_lewiner_f32.restype = ctypes.c_long
        _lib.mc_lewiner_f32.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_double, ctypes.c_void_p, ctypes.c_long,
                                        ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_long)]
        _lib.mc_select_tiling.restype = ctypes.c_int
        _lib.mc_select_tiling.argtypes = [ctypes.c_void_p]
    return _lib


class MCValueError(ValueError):
    pass


class MCRuntimeError(RuntimeError):
    pass


def marching_cubes_soup(volume, level=0.0, return_active=False):
    """Stand-in for core._marching_cubes (core.py:16-18): (3T,3) float32 soup in
    array-axis (voxel index) coordinates; raises like skimage does."""
    vol = np.ascontiguousarray(volume, np.float32)
    if vol.ndim != 3:
        raise MCValueError('Input volume should be a 3D numpy array.')
    lib = mc_lib()
    status = ctypes.c_int(0)
    active = ctypes.c_long(0)
    n0, n1, n2 = vol.shape
    cap = max(1, (n0 - 1) * (n1 - 1) * (n2 - 1)) * 12
    out = np.empty((min(cap, 1 << 24), 9), np.float32)
    n = lib.mc_lewiner_f32(vol.ctypes.data, n0, n1, n2, float(level), out.ctypes.data,
                           out.shape[0], ctypes.byref(status), ctypes.byref(active))
    if n < 0:
        out = np.empty((cap, 9), np.float32)
        n = lib.mc_lewiner_f32(vol.ctypes.data, n0, n1, n2, float(level), out.ctypes.data,
                               out.shape[0], ctypes.byref(status), ctypes.byref(active))
    if status.value == 1:
        raise MCValueError('Surface level must be within volume data range / array at least 2x2x2.')
    if status.value == 2:
        raise MCRuntimeError('No surface found at the given iso value.')
    soup = out[:n].reshape(-1, 3).copy()
    return (soup, active.value) if return_active else soup


# --------------------------------------------------------------------------
# engine (core.py)

def cartesian_product(*arrays):             # core.py:20-26
    grids = np.meshgrid(*arrays, indexing='ij')
    return np.stack([g.reshape(-1) for g in grids], axis=-1)


def skip(sdf, job):                         # core.py:28-43
    X, Y, Z = job
    x0, x1 = X[0], X[-1]
    y0, y1 = Y[0], Y[-1]
    z0, z1 = Z[0], Z[-1]
    x = (x0 + x1) / 2
    y = (y0 + y1) / 2
    z = (z0 + z1) / 2
    r = abs(evaluate(sdf, np.array([(x, y, z)])).reshape(-1)[0])
    d = np.linalg.norm(np.array((x - x0, y - y0, z - z0)))
    if r <= d:
        return False
    corners = np.array(list(itertools.product((x0, x1), (y0, y1), (z0, z1))))
    values = evaluate(sdf, corners).reshape(-1)
    return bool(np.all(values > 0) if values[0] > 0 else np.all(values < 0))


def worker(sdf, job, sparse=True):          # core.py:45-60
    X, Y, Z = job
    if sparse and skip(sdf, job):
        return None
    P = cartesian_product(X, Y, Z)
    volume = evaluate(sdf, P).reshape((len(X), len(Y), len(Z)))
    try:
        points = marching_cubes_soup(volume)
    except Exception:
        return []
    scale = np.array([X[1] - X[0], Y[1] - Y[0], Z[1] - Z[0]])
    offset = np.array([X[0], Y[0], Z[0]])
    return points * scale + offset


def estimate_bounds(sdf):                   # core.py:62-82
    s = 16
    x0 = y0 = z0 = -1e9
    x1 = y1 = z1 = 1e9
    prev = None
    for _ in range(32):
        X = np.linspace(x0, x1, s)
        Y = np.linspace(y0, y1, s)
        Z = np.linspace(z0, z1, s)
        d = np.array([X[1] - X[0], Y[1] - Y[0], Z[1] - Z[0]])
        threshold = np.linalg.norm(d) / 2
        if threshold == prev:
            break
        prev = threshold
        P = cartesian_product(X, Y, Z)
        volume = evaluate(sdf, P).reshape((len(X), len(Y), len(Z)))
        where = np.argwhere(np.abs(volume) <= threshold)
        x1, y1, z1 = (x0, y0, z0) + where.max(axis=0) * d + d / 2
        x0, y0, z0 = (x0, y0, z0) + where.min(axis=0) * d - d / 2
    return ((x0, y0, z0), (x1, y1, z1))


def lattice(bounds, step=None, samples=2 ** 22):   # core.py:94-112
    (x0, y0, z0), (x1, y1, z1) = bounds
    if step is None and samples is not None:
        volume = (x1 - x0) * (y1 - y0) * (z1 - z0)
        step = (volume / samples) ** (1 / 3)
    try:
        dx, dy, dz = step
    except TypeError:
        dx = dy = dz = step
    return np.arange(x0, x1, dx), np.arange(y0, y1, dy), np.arange(z0, z1, dz)


def batches_of(X, Y, Z, batch_size=32):     # core.py:114-119
    s = batch_size
    Xs = [X[i:i + s + 1] for i in range(0, len(X), s)]
    Ys = [Y[i:i + s + 1] for i in range(0, len(Y), s)]
    Zs = [Z[i:i + s + 1] for i in range(0, len(Z), s)]
    return list(itertools.product(Xs, Ys, Zs))


def generate(sdf, step=None, bounds=None, samples=2 ** 22, batch_size=32, sparse=True,
             batch_subset=None, workers=1):
    """core.generate (core.py:84-150) without the printing.

    Returns (points (3T,3) float64, stats dict).  ``batch_subset`` (iterable of
    batch indices) restricts the work for bounded-time baselines; ``workers`` > 1
    uses a ThreadPool over batches exactly like the reference (core.py:131-133).
    """
    if bounds is None:
        bounds = estimate_bounds(sdf)
    X, Y, Z = lattice(bounds, step, samples)
    batches = batches_of(X, Y, Z, batch_size)
    if batch_subset is not None:
        batches = [batches[i] for i in batch_subset]
    pts = []
    skipped = empty = nonempty = 0
    evaluated = 0
    if workers > 1:
        from multiprocessing.pool import ThreadPool
        with ThreadPool(workers) as pool:
            results = list(pool.imap(lambda job: worker(sdf, job, sparse), batches))
    else:
        results = (worker(sdf, job, sparse) for job in batches)
    for job, res in zip(batches, results):
        if res is None:
            skipped += 1
            continue
        evaluated += len(job[0]) * len(job[1]) * len(job[2])
        if len(res) == 0:
            empty += 1
        else:
            nonempty += 1
            pts.append(res)
    points = np.concatenate(pts) if pts else np.zeros((0, 3))
    stats = dict(skipped=skipped, empty=empty, nonempty=nonempty, batches=len(batches),
                 samples=sum(len(a) * len(b) * len(c) for a, b, c in batches),
                 evaluated=evaluated, triangles=len(points) // 3,
                 dims=(len(X), len(Y), len(Z)))
    return points, stats

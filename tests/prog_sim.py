"""numpy float64 simulator of the evaluator's linear program (sdf_b200/ir.py machine model).

TEST INFRASTRUCTURE: restates, op by op, what csrc/sdf_ops.cuh does on the device (accurate / float64
path), so that the lowering (`ir.lower`: stack depths, point slots, fold order, constant layout) can be
checked against the oracle on a machine without a GPU.  Vectorised over the sample points.
"""
import numpy as np

from sdf_b200 import ir


def _len(*v):
    return np.sqrt(sum(c * c for c in v))


def _clip01(v):
    return np.minimum(np.maximum(v, 0.0), 1.0)


def _pymod(a, b):
    return np.mod(a, b)


def _sign(v):
    return np.sign(v)


def _bounce_out(t):
    a = (121 * t * t) / 16
    b = (363 / 40.0 * t * t) - (99 / 10.0 * t) + 17 / 5.0
    c = (4356 / 361.0 * t * t) - (35442 / 1805.0 * t) + 16061 / 1805.0
    d = (54 / 5.0 * t * t) - (513 / 25.0 * t) + 268 / 25.0
    return np.where(t < 4 / 11.0, a, np.where(t < 8 / 11.0, b, np.where(t < 9 / 10.0, c, d)))


def ease_eval(i, t):
    pi = np.pi
    with np.errstate(all='ignore'):
        if i == 0: return t
        if i == 1: return t * t
        if i == 2: return -t * (t - 2)
        if i == 3:
            u = 2 * t - 1
            return np.where(t < 0.5, 2 * t * t, -0.5 * (u * (u - 2) - 1))
        if i == 4: return t * t * t
        if i == 5:
            u = t - 1
            return u * u * u + 1
        if i == 6:
            u = t * 2; v = u - 2
            return np.where(u < 1, 0.5 * u * u * u, 0.5 * (v * v * v + 2))
        if i == 7: return t * t * t * t
        if i == 8:
            u = t - 1
            return -(u * u * u * u - 1)
        if i == 9:
            u = t * 2; v = u - 2
            return np.where(u < 1, 0.5 * u * u * u * u, -0.5 * (v * v * v * v - 2))
        if i == 10: return t * t * t * t * t
        if i == 11:
            u = t - 1
            return u * u * u * u * u + 1
        if i == 12:
            u = t * 2; v = u - 2
            return np.where(u < 1, 0.5 * u ** 5, 0.5 * (v ** 5 + 2))
        if i == 13: return -np.cos(t * pi / 2) + 1
        if i == 14: return np.sin(t * pi / 2)
        if i == 15: return -0.5 * (np.cos(pi * t) - 1)
        if i == 16: return np.where(t == 0, 0.0, np.exp2(10 * (t - 1)))
        if i == 17: return np.where(t == 1, 1.0, 1 - np.exp2(-10 * t))
        if i == 18:
            return np.where(t == 0, 0.0, np.where(t == 1, 1.0, np.where(t < 0.5, 0.5 * np.exp2(20 * t - 10),
                                                                       1 - 0.5 * np.exp2(-20 * t + 10))))
        if i == 19: return -1 * (np.sqrt(1 - t * t) - 1)
        if i == 20:
            u = t - 1
            return np.sqrt(1 - u * u)
        if i == 21:
            u = t * 2; v = u - 2
            return np.where(u < 1, -0.5 * (np.sqrt(1 - u * u) - 1), 0.5 * (np.sqrt(1 - v * v) + 1))
        if i == 22:
            u = t - 1; k = 0.5
            return -1 * (np.exp2(10 * u) * np.sin((u - k / 4) * (2 * pi) / k))
        if i == 23:
            k = 0.5
            return np.exp2(-10 * t) * np.sin((t - k / 4) * (2 * pi / k)) + 1
        if i == 24:
            u = t * 2; v = u - 1; k = 0.5
            s = np.sin((v - k / 4) * 2 * pi / k)
            return np.where(u < 1, -0.5 * (np.exp2(10 * v) * s), np.exp2(-10 * v) * s * 0.5 + 1)
        if i == 25:
            k = 1.70158
            return t * t * ((k + 1) * t - k)
        if i == 26:
            k = 1.70158; u = t - 1
            return u * u * ((k + 1) * u + k) + 1
        if i == 27:
            k = 1.70158 * 1.525; u = t * 2; v = u - 2
            return np.where(u < 1, 0.5 * (u * u * ((k + 1) * u - k)), 0.5 * (v * v * ((k + 1) * v + k) + 2))
        if i == 28: return 1 - _bounce_out(1 - t)
        if i == 29: return _bounce_out(t)
        if i == 30:
            return np.where(t < 0.5, (1 - _bounce_out(1 - 2 * t)) * 0.5, _bounce_out(2 * t - 1) * 0.5 + 0.5)
        if i == 31: return np.where(t < 1, 0.0, 1.0)
        if i == 32: return np.where(t > 0, 1.0, 0.0)
        if i == 33: return np.where(t < 0.5, 0.0, 1.0)
    raise ValueError(i)


def run(program, P):
    """Evaluate the lowered `program` (ir.Program) at points P (n,3); returns d (n,)."""
    code = np.asarray(program.code, np.uint32)
    C = np.asarray(program.consts, np.float64)
    P = np.asarray(P, np.float64)
    x, y, z = P[:, 0].copy(), P[:, 1].copy(), P[:, 2].copy()
    n = len(P)
    d = np.zeros(n)
    s0 = np.zeros(n)
    vs = [np.zeros(n) for _ in range(ir.MAX_VSTACK + 1)]
    ps = [[np.zeros(n) for _ in range(3)] for _ in range(ir.MAX_PSLOTS)]
    mx, mn = np.maximum, np.minimum

    def push(a):
        if a > 0:
            vs[a - 1] = s0

    with np.errstate(all='ignore'):
        for w in code:
            w = int(w)
            name, a, c = ir.OPNAME[w & 0xff], (w >> 8) & 0xff, C[(w >> 16):]
            if name == 'SPHERE':
                d = _len(x - c[0], y - c[1], z - c[2]) - c[3]
            elif name == 'PLANE':
                d = (c[3] - x) * c[0] + (c[4] - y) * c[1] + (c[5] - z) * c[2]
            elif name in ('BOX', 'ROUNDED_BOX'):
                if name == 'BOX':
                    qx, qy, qz = abs(x - c[0]) - c[3], abs(y - c[1]) - c[4], abs(z - c[2]) - c[5]; r = 0.0
                else:
                    r = c[3]
                    qx, qy, qz = abs(x) - c[0] + r, abs(y) - c[1] + r, abs(z) - c[2] + r
                d = _len(mx(qx, 0), mx(qy, 0), mx(qz, 0)) + mn(mx(qx, mx(qy, qz)), 0) - r
            elif name == 'WIREFRAME_BOX':
                ht = c[3] / 2
                px, py, pz = abs(x) - c[0] - ht, abs(y) - c[1] - ht, abs(z) - c[2] - ht
                qx, qy, qz = abs(px + ht) - ht, abs(py + ht) - ht, abs(pz + ht) - ht
                g = lambda u, v, t: _len(mx(u, 0), mx(v, 0), mx(t, 0)) + mn(mx(u, mx(v, t)), 0)
                d = mn(mn(g(px, qy, qz), g(qx, py, qz)), g(qx, qy, pz))
            elif name == 'TORUS':
                d = _len(_len(x, y) - c[0], z) - c[1]
            elif name == 'CAPSULE':
                ax, ay, az, r = c[0], c[1], c[2], c[6]
                bx, by, bz = c[3] - ax, c[4] - ay, c[5] - az
                px, py, pz = x - ax, y - ay, z - az
                h = _clip01((px * bx + py * by + pz * bz) / (bx * bx + by * by + bz * bz))
                d = _len(px - bx * h, py - by * h, pz - bz * h) - r
            elif name == 'CYLINDER':
                d = _len(x, y) - c[0]
            elif name == 'CAPPED_CYLINDER':
                ax, ay, az, r = c[0], c[1], c[2], c[6]
                bx, by, bz = c[3] - ax, c[4] - ay, c[5] - az
                baba = bx * bx + by * by + bz * bz
                px, py, pz = x - ax, y - ay, z - az
                paba = px * bx + py * by + pz * bz
                xx = _len(px * baba - bx * paba, py * baba - by * paba, pz * baba - bz * paba) - r * baba
                yy = abs(paba - baba * 0.5) - baba * 0.5
                x2, y2 = xx * xx, yy * yy * baba
                dd = np.where(mx(xx, yy) < 0, -mn(x2, y2), np.where(xx > 0, x2, 0) + np.where(yy > 0, y2, 0))
                d = _sign(dd) * np.sqrt(abs(dd)) / baba
            elif name == 'ROUNDED_CYLINDER':
                ra, rb, hh = c[0], c[1], c[2] / 2
                d0 = _len(x, y) - ra + rb
                d1 = abs(z) - hh + rb
                d = mn(mx(d0, d1), 0) + _len(mx(d0, 0), mx(d1, 0)) - rb
            elif name == 'CAPPED_CONE':
                ax, ay, az, ra, rb = c[0], c[1], c[2], c[6], c[7]
                bx, by, bz = c[3] - ax, c[4] - ay, c[5] - az
                rba = rb - ra
                baba = bx * bx + by * by + bz * bz
                k = rba * rba + baba
                px, py, pz = x - ax, y - ay, z - az
                papa = px * px + py * py + pz * pz
                paba = (px * bx + py * by + pz * bz) / baba
                xx = np.sqrt(papa - paba * paba * baba)
                cax = mx(0, xx - np.where(paba < 0.5, ra, rb))
                cay = abs(paba - 0.5) - 0.5
                f = _clip01((rba * (xx - ra) + paba * baba) / k)
                cbx = xx - ra - f * rba
                cby = paba - f
                s = np.where((cbx < 0) & (cay < 0), -1.0, 1.0)
                d = s * np.sqrt(mn(cax * cax + cay * cay * baba, cbx * cbx + cby * cby * baba))
            elif name == 'ROUNDED_CONE':
                r1, r2, h = c[0], c[1], c[2]
                b = (r1 - r2) / h
                aa = np.sqrt(1 - b * b)
                qx, qy = _len(x, y), z
                k = qx * (-b) + qy * aa
                c1 = _len(qx, qy) - r1
                c2 = _len(qx, qy - h) - r2
                c3 = (qx * aa + qy * b) - r1
                d = np.where(k < 0, c1, np.where(k > aa * h, c2, c3))
            elif name == 'ELLIPSOID':
                sx, sy, sz = c[0], c[1], c[2]
                k0 = _len(x / sx, y / sy, z / sz)
                k1 = _len(x / (sx * sx), y / (sy * sy), z / (sz * sz))
                d = k0 * (k0 - 1) / k1
            elif name == 'PYRAMID':
                h = c[0]
                m2 = h * h + 0.25
                a0, a1 = abs(x) - 0.5, abs(y) - 0.5
                sw = a1 > a0
                a0, a1 = np.where(sw, a1, a0), np.where(sw, a0, a1)
                px, py, pz = a0, z, a1
                qx, qy, qz = pz, h * py - 0.5 * px, h * px + 0.5 * py
                s = mx(-qx, 0)
                t = _clip01((qy - 0.5 * pz) / (m2 + 0.25))
                aa = m2 * (qx + s) * (qx + s) + qy * qy
                bb = m2 * (qx + 0.5 * t) * (qx + 0.5 * t) + (qy - m2 * t) * (qy - m2 * t)
                d2 = np.where(mn(qy, -qx * m2 - qy * 0.5) > 0, 0.0, mn(aa, bb))
                d = np.sqrt((d2 + qz * qz) / m2) * _sign(mx(qz, -py))
            elif name == 'TETRAHEDRON':
                d = (mx(abs(x + y) - z, abs(x - y) + z) - c[0]) / np.sqrt(3.0)
            elif name == 'OCTAHEDRON':
                d = (abs(x) + abs(y) + abs(z) - c[0]) * c[1]
            elif name in ('DODECAHEDRON', 'ICOSAHEDRON'):
                r, vx, vy, vz = c[0], c[1], c[2], c[3]
                px, py, pz = abs(x / r), abs(y / r), abs(z / r)
                aa = px * vx + py * vy + pz * vz
                bb = px * vz + py * vx + pz * vy
                cc = px * vy + py * vz + pz * vx
                if name == 'DODECAHEDRON':
                    d = (mx(mx(aa, bb), cc) - vx) * r
                else:
                    dd = (px * c[4] + py * c[4] + pz * c[4]) - vx
                    d = mx(mx(mx(aa, bb), cc) - vx, dd) * r
            elif name == 'TRANSLATE':
                x, y, z = x - c[0], y - c[1], z - c[2]
            elif name == 'SCALE_P':
                x, y, z = x / c[0], y / c[1], z / c[2]
            elif name == 'ROTATE':
                x, y, z = (x * c[0] + y * c[3] + z * c[6], x * c[1] + y * c[4] + z * c[7], x * c[2] + y * c[5] + z * c[8])
            elif name == 'TWIST':
                sn, cs = np.sin(c[0] * z), np.cos(c[0] * z)
                x, y = cs * x - sn * y, sn * x + cs * y
            elif name == 'BEND':
                sn, cs = np.sin(c[0] * x), np.cos(c[0] * x)
                x, y = cs * x - sn * y, sn * x + cs * y
            elif name == 'ELONGATE':
                push(a)
                qx, qy, qz = abs(x) - c[0], abs(y) - c[1], abs(z) - c[2]
                s0 = mn(mx(qx, mx(qy, qz)), 0)
                x, y, z = mx(qx, 0), mx(qy, 0), mx(qz, 0)
            elif name == 'CIRCULAR_PRE':
                da = c[0]
                r = np.hypot(x, y)
                ang = _pymod(np.arctan2(y, x), da)
                ps[a] = [np.cos(ang) * r, np.sin(ang) * r, z]
                x, y = np.cos(ang - da) * r, np.sin(ang - da) * r
            elif name == 'REPEAT_CELL':
                out = []
                for k, p in enumerate((x, y, z)):
                    sp = c[k]
                    q = p / sp if sp != 0 else np.zeros_like(p)
                    idx = np.rint(q)
                    if c[6] != 0:
                        idx = mn(mx(idx, -c[3 + k]), c[3 + k])
                    out.append(p - sp * (idx + c[7 + k]))
                x, y, z = out
            elif name == 'MULC':
                d = d * c[0]
            elif name == 'NEGATE':
                d = -d
            elif name == 'ADDC':
                d = d + c[0]
            elif name == 'SHELL':
                d = abs(d) - c[0] / 2
            elif name in ('ADD_POP', 'UNION', 'UNION_K', 'DIFFERENCE', 'DIFFERENCE_K', 'INTERSECTION', 'INTERSECTION_K', 'BLEND',
                          'EXTRUDE_POST', 'SLICE_SELECT'):
                d1, d2 = s0, d
                if name == 'ADD_POP': d = d2 + d1
                elif name == 'UNION': d = mn(d1, d2)
                elif name == 'DIFFERENCE': d = mx(d1, -d2)
                elif name == 'INTERSECTION': d = mx(d1, d2)
                elif name == 'BLEND': d = c[0] * d2 + (1 - c[0]) * d1
                elif name == 'UNION_K':
                    h = _clip01(0.5 + 0.5 * (d2 - d1) / c[0])
                    d = d2 + (d1 - d2) * h - c[0] * h * (1 - h)
                elif name == 'DIFFERENCE_K':
                    h = _clip01(0.5 - 0.5 * (d2 + d1) / c[0])
                    d = d1 + (-d2 - d1) * h + c[0] * h * (1 - h)
                elif name == 'INTERSECTION_K':
                    h = _clip01(0.5 - 0.5 * (d2 - d1) / c[0])
                    d = d2 + (d1 - d2) * h + c[0] * h * (1 - h)
                elif name == 'EXTRUDE_POST':
                    d = mn(mx(d2, d1), 0) + _len(mx(d2, 0), mx(d1, 0))
                elif name == 'SLICE_SELECT':
                    d = np.where(d1 <= 0, -d2, d1)
                if a > 0:
                    s0 = vs[a - 1]
            elif name == 'CIRCLE':
                d = _len(x - c[0], y - c[1]) - c[2]
            elif name == 'LINE2':
                d = (c[2] - x) * c[0] + (c[3] - y) * c[1]
            elif name == 'RECTANGLE':
                qx, qy = abs(x - c[0]) - c[2], abs(y - c[1]) - c[3]
                d = _len(mx(qx, 0), mx(qy, 0)) + mn(mx(qx, qy), 0)
            elif name == 'ROUNDED_RECTANGLE':
                xp, yp = x > 0, y > 0
                r = np.where(xp, np.where(yp, c[2], c[3]), np.where(yp, c[5], c[4]))
                qx, qy = abs(x) - c[0] + r, abs(y) - c[1] + r
                d = mn(mx(qx, qy), 0) + _len(mx(qx, 0), mx(qy, 0)) - r
            elif name == 'EQUILATERAL_TRIANGLE':
                k = np.sqrt(3.0)
                px, py = abs(x) - 1, y + 1 / k
                fold = px + k * py > 0
                px, py = np.where(fold, (px - k * py) / 2, px), np.where(fold, (-k * px - py) / 2, py)
                px = px - mn(mx(px, -2), 0)
                d = -_len(px, py) * _sign(py)
            elif name == 'HEXAGON':
                r = c[0]
                k0, k1, k2 = np.sqrt(3.0) / -2, 0.5, 0.57735026918962573
                px, py = abs(x), abs(y)
                h = mn(k0 * px + k1 * py, 0)
                px, py = px - 2 * k0 * h, py - 2 * k1 * h
                px = px - mn(mx(px, -k2 * r), k2 * r)
                py = py - r
                d = _len(px, py) * _sign(py)
            elif name == 'ROUNDED_X':
                px, py = abs(x), abs(y)
                q = mn(px + py, c[0]) * 0.5
                d = _len(px - q, py - q) - c[1]
            elif name == 'POLYGON':
                m = a
                dd = (x - c[0]) ** 2 + (y - c[1]) ** 2
                s = np.ones(n)
                for q in range(m):
                    j = (q + m - 1) % m
                    vix, viy, vjx, vjy = c[2 * q], c[2 * q + 1], c[2 * j], c[2 * j + 1]
                    ex, ey = vjx - vix, vjy - viy
                    wx, wy = x - vix, y - viy
                    h = _clip01((wx * ex + wy * ey) / (ex * ex + ey * ey))
                    bx, by = wx - ex * h, wy - ey * h
                    dd = mn(dd, bx * bx + by * by)
                    c1, c2, c3 = y >= viy, y < vjy, ex * wy > ey * wx
                    s = np.where((c1 & c2 & c3) | (~c1 & ~c2 & ~c3), -s, s)
                d = s * np.sqrt(dd)
            elif name == 'VESICA':
                r, dd = c[0], c[1]
                b = np.sqrt(r * r - dd * dd)
                px, py = abs(x), abs(y)
                d = np.where((py - b) * dd > px * b, _len(px, py - b), _len(px + dd, py) - r)
            elif name == 'REVOLVE':
                x, y = _len(x, y) - c[0], z
            elif name == 'EXTRUDE_PRE':
                push(a)
                s0 = abs(z) - c[0] / 2
            elif name == 'SLICE_PRE':
                z = np.zeros(n)
            elif name == 'TPARAM_LINEAR':
                ax, ay, az = c[3], c[4], c[5]
                d = _clip01(((x - c[0]) * ax + (y - c[1]) * ay + (z - c[2]) * az) / (ax * ax + ay * ay + az * az))
            elif name == 'TPARAM_RADIAL':
                d = _clip01((np.hypot(x, y) - c[0]) / (c[1] - c[0]))
            elif name == 'TPARAM_Z':
                d = mn(mx(z / c[0], -0.5), 0.5) + 0.5
            elif name == 'EASE':
                if a != 0:
                    d = ease_eval(a, d)
            elif name == 'DISPLACE':
                x, y, z = x + d * c[0], y + d * c[1], z + d * c[2]
            elif name in ('LERP_POP2', 'MIX_POP2'):
                t, d2, d1 = d, s0, vs[a]
                d = t * d2 + (1 - t) * d1 if name == 'LERP_POP2' else d1 + (d2 - d1) * t
                if a > 0:
                    s0 = vs[a - 1]
            elif name == 'WRAP_PRE':
                dd = np.hypot(x, y) - c[0]
                d = (np.arctan2(y, x) + np.pi) / (2 * np.pi)
                y = -dd
            elif name == 'WRAP_POST':
                x = c[0] + (c[1] - c[0]) * d
            elif name == 'PUSH':
                push(a)
                s0 = d
            elif name == 'SAVEP':
                ps[a] = [x, y, z]
            elif name == 'LOADP':
                x, y, z = ps[a]
            else:
                raise ValueError('unknown op ' + name)
    return d

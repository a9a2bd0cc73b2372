"""A-priori error bound of the fast float32 classification pass (ir.py machine model, csrc/sdf_ops.cuh FAST path).

Precision `exact` classifies every lattice sample with the fast float32 evaluator and re-evaluates in float64
whatever lies within a guard band ``tau`` of the surface (plus every corner of a candidate cube).  The mesh is
the reference's only if the fast pass never errs by more than ``tau`` -- a property of the TREE, not of a test
corpus: a warp that divides by a small constant difference (``bend_radial`` with r1 - r0 = 1e-3), a steep
easing, a large ``twist`` rate or a tiny scale factor amplifies the float32 rounding of the coordinates far
beyond the usual few ulps.

:func:`error_bound` pushes (magnitude, absolute error) bounds through the lowered program, one rule per opcode:

    error_out = sum_i |d f / d in_i|_max * error_in_i  +  c_op * u * magnitude_out     (u = 2^-23)

with the Lipschitz factors of the node library (an SDF primitive is 1-Lipschitz in the point, a hard fold takes
the larger error, a smooth fold at most the sum, ``t = (r - r0) / (r1 - r0)`` scales by 1 / |r1 - r0|, an easing
by its steepest slope ...).  Rules are deliberately coarse and conservative (whole-lattice magnitudes, no
cancellation credit).  An opcode whose slope is unbounded (the square easings, ``in_circ`` at its end point)
returns ``inf``: every sample is then re-evaluated in float64 -- still exact, merely without the fast pass's
benefit.

The absolute bound is far too pessimistic to be the guard band itself (worst-case magnitudes everywhere); what
it isolates reliably is AMPLIFICATION.  :func:`gain` runs the analysis twice -- with the tree's constants, and
with every amplifying factor clamped to one (the kind of tree the corpus-validated heuristic of
``core.mixed_tau`` was tuned on) -- and returns the ratio (>= 1).  ``core.mixed_tau`` multiplies its heuristic
by it: unit-gain trees (all five BASELINE configurations) keep their guard band, a tree that divides by
r1 - r0 = 1e-3 gets a band a thousand times wider, a tree with a jump gets float64 everywhere.
"""
import math

import numpy as np

from .ir import NCONST, OPNAME

U = 2.0 ** -23            # float32 unit roundoff (relative spacing; a rounded operation errs by <= 0.5 ulp, the FAST
#                           division / square root by <= 2 ulp -- the rule constants c_op count in units of U)
TRIG_ABS = 5.0e-7         # MUFU sine / cosine after one Cody-Waite step, polynomial atan2: absolute error
SQRT2, SQRT3 = math.sqrt(2.0), math.sqrt(3.0)
INF = float('inf')

# steepest slope of each easing on [0, 1] (ease.py:3-162), in the order of sdf_ops.cuh ease_eval(); inf = a jump or
# an unbounded derivative
EASE_SLOPE = [1.0, 2.0, 2.0, 2.0, 3.0, 3.0, 3.0, 4.0, 4.0, 4.0, 5.0, 5.0, 5.0,
              math.pi / 2, math.pi / 2, math.pi / 2,
              10 * math.log(2), 10 * math.log(2), 10 * math.log(2),
              INF, INF, INF,                       # circ: sqrt(1 - t^2) is vertical at an end point
              100.0, 100.0, 200.0,                 # elastic: 2^(10 u) * sin(...) with period 0.5 (slope <= 2 pi / 0.5 * 2 + 10 ln 2)
              4.5, 4.5, 6.0,                       # back
              8.0, 8.0, 16.0,                      # bounce: piecewise quadratics, |slope| <= 2 * 121 / 16 * (4 / 11)... <= 8
              INF, INF, INF]                       # square: jumps

# c_op of the distance primitives: operations along the longest dependency chain (approximate div / sqrt = 2)
PRIM_COST = {
    'SPHERE': 8, 'PLANE': 8, 'BOX': 10, 'ROUNDED_BOX': 12, 'WIREFRAME_BOX': 16, 'TORUS': 12, 'CAPSULE': 16,
    'CYLINDER': 6, 'CAPPED_CYLINDER': 40, 'ROUNDED_CYLINDER': 14, 'CAPPED_CONE': 48, 'ROUNDED_CONE': 32,
    'ELLIPSOID': 32, 'PYRAMID': 40, 'TETRAHEDRON': 8, 'OCTAHEDRON': 8, 'DODECAHEDRON': 12, 'ICOSAHEDRON': 12,
    'CIRCLE': 6, 'LINE2': 6, 'RECTANGLE': 10, 'ROUNDED_RECTANGLE': 12, 'EQUILATERAL_TRIANGLE': 20, 'HEXAGON': 16,
    'ROUNDED_X': 10, 'VESICA': 14,
}
# primitives whose closed form is NOT 1-Lipschitz / amplifies (d3.py:252-259 ellipsoid: k0 (k0 - 1) / k1 with
# k1 -> 0 at the centre; capped cone / cylinder / pyramid: differences of squares under a root)
PRIM_GAIN = {'ELLIPSOID': 4.0, 'CAPPED_CONE': 4.0, 'CAPPED_CYLINDER': 2.0, 'PYRAMID': 2.0, 'ROUNDED_CONE': 2.0}


class _V:
    """(magnitude bound, absolute error bound) of one machine register."""
    __slots__ = ('m', 'e')

    def __init__(self, m=0.0, e=0.0):
        self.m, self.e = float(m), float(e)

    def copy(self):
        return _V(self.m, self.e)


def error_bound(program, coord_max, unit_gain=False):
    """Absolute error bound of the FAST float32 evaluation of ``program`` at any point with |x|,|y|,|z| <=
    ``coord_max``.  inf when some node has no finite slope.  ``unit_gain``: every amplifying factor (reciprocals of
    small constants, slopes, rates, scale factors) clamped to <= 1 -- the reference analysis of :func:`gain`."""
    code, consts = program.code, program.consts

    def amp(v):                               # an amplification factor as the analysis uses it
        return min(v, 1.0) if unit_gain else v
    # p.m: bound on |x|, |y|, |z|; p.e: bound on the EUCLIDEAN error of the point (rotations keep it)
    p = _V(coord_max, SQRT3 * U * coord_max)  # each coordinate: the float64 axis value rounded to float32
    d = _V()
    s0 = _V()
    vs = {}
    ps = {}
    for w in code:
        w = int(w)
        name, a, ci = OPNAME[w & 0xff], (w >> 8) & 0xff, w >> 16
        n = NCONST[name] if NCONST[name] >= 0 else 2 * a
        c = [float(v) for v in consts[ci:ci + n]]
        cm = max([abs(v) for v in c if math.isfinite(v)] + [0.0])
        if name in PRIM_COST or name == 'POLYGON':
            cost = PRIM_COST.get(name, 10 * max(a, 1))
            gain = amp(PRIM_GAIN.get(name, 1.0))
            m = SQRT3 * p.m + cm
            d = _V(m, gain * p.e + cost * U * m)
        elif name == 'TRANSLATE':
            p = _V(p.m + cm, p.e + SQRT3 * U * (p.m + cm))
        elif name == 'SCALE_P':
            smin = min(abs(v) for v in c)
            if smin == 0.0:
                return INF
            inv = 1.0 if unit_gain else 1.0 / smin          # (a scale is undone by its MULC: never an amplifier by itself)
            p = _V(p.m * inv, p.e * inv + 2 * SQRT3 * U * p.m * inv)
        elif name == 'ROTATE':
            M = np.array(c, dtype=np.float64).reshape(3, 3)                 # P' = P @ M: a rotation (spectral norm 1)
            sn = amp(float(np.linalg.norm(M, 2)))
            col = max(abs(c[0 + k]) + abs(c[3 + k]) + abs(c[6 + k]) for k in range(3))
            p = _V(col * p.m, sn * p.e + 3 * SQRT3 * U * col * p.m)
        elif name in ('TWIST', 'BEND'):
            k = amp(abs(c[0]))
            ea = k * p.e + 2 * U * k * p.m + TRIG_ABS
            p = _V(SQRT2 * p.m, p.e + SQRT2 * p.m * ea + 3 * SQRT3 * U * SQRT2 * p.m)
        elif name == 'ELONGATE':
            vs_push(vs, a, s0)
            s0 = _V(p.m + cm, p.e + U * (p.m + cm))
            p = _V(p.m + cm, p.e + U * (p.m + cm))
        elif name == 'CIRCULAR_PRE':
            ea = 2 * TRIG_ABS + 4 * U * math.pi
            q = _V(SQRT2 * p.m, p.e + SQRT2 * p.m * ea + 6 * SQRT3 * U * SQRT2 * p.m)
            ps[a] = q.copy()
            p = q
        elif name == 'REPEAT_CELL':
            p = _V(p.m + cm, p.e + 3 * U * (p.m + cm))
        elif name in ('MULC',):
            mk = 1.0 if unit_gain else abs(c[0])
            d = _V(mk * d.m, mk * d.e + U * mk * d.m)
        elif name == 'NEGATE':
            pass
        elif name in ('ADDC', 'SHELL'):
            d = _V(d.m + cm, d.e + U * (d.m + cm))
        elif name == 'ADD_POP':
            d = _V(d.m + s0.m, d.e + s0.e + U * (d.m + s0.m))
            s0 = vs_pop(vs, a)
        elif name in ('UNION', 'DIFFERENCE', 'INTERSECTION', 'SLICE_SELECT'):
            d = _V(max(d.m, s0.m), max(d.e, s0.e))
            s0 = vs_pop(vs, a)
        elif name in ('UNION_K', 'DIFFERENCE_K', 'INTERSECTION_K'):
            K = abs(c[0])
            if K == 0.0:
                return INF
            m = max(d.m, s0.m) + K
            eh = (d.e + s0.e) / (2 * K) + 4 * U
            d = _V(m, max(d.e, s0.e) + K * eh * eh + 8 * U * m)       # a convex combination; K h (1 - h) is stationary in h
            s0 = vs_pop(vs, a)
        elif name == 'BLEND':
            K = c[0]
            d = _V(abs(K) * d.m + abs(1 - K) * s0.m, abs(K) * d.e + abs(1 - K) * s0.e + 3 * U * (d.m + s0.m))
            s0 = vs_pop(vs, a)
        elif name == 'REVOLVE':
            m = SQRT2 * p.m + cm
            p = _V(m, p.e + 4 * U * m)
        elif name == 'EXTRUDE_PRE':
            vs_push(vs, a, s0)
            s0 = _V(p.m + cm, p.e + U * (p.m + cm))
        elif name == 'EXTRUDE_POST':
            m = SQRT2 * max(d.m, s0.m)
            d = _V(m, d.e + s0.e + 6 * U * m)
            s0 = vs_pop(vs, a)
        elif name == 'SLICE_PRE':
            pass
        elif name == 'TPARAM_LINEAR':
            ab = math.sqrt(c[3] ** 2 + c[4] ** 2 + c[5] ** 2)
            if ab == 0.0:
                return INF
            d = _V(1.0, (p.e + 6 * U * (p.m + cm)) * amp(1.0 / ab) + 2 * U)
        elif name == 'TPARAM_RADIAL':
            dr = abs(c[1] - c[0])
            if dr == 0.0:
                return INF
            d = _V(1.0, (p.e + 4 * U * (SQRT2 * p.m + cm)) * amp(1.0 / dr) + 2 * U)
        elif name == 'TPARAM_Z':
            if c[0] == 0.0:
                return INF
            d = _V(1.0, (p.e + 2 * U * p.m) * amp(1.0 / abs(c[0])) + 2 * U)
        elif name == 'EASE':
            if a:
                slope = EASE_SLOPE[a] if a < len(EASE_SLOPE) else INF
                if not math.isfinite(slope):
                    if not unit_gain:
                        return INF
                    slope = 1.0
                slope = amp(slope)
                d = _V(max(1.0, d.m) * 1.7, slope * d.e + 8 * U * 1.7 + (TRIG_ABS if 13 <= a <= 15 or 22 <= a <= 24 else 0.0))
        elif name == 'DISPLACE':
            v = cm
            p = _V(p.m + v * d.m, p.e + amp(v) * d.e + 2 * U * (p.m + v * d.m))
        elif name in ('LERP_POP2', 'MIX_POP2'):
            d1 = vs.get(a, _V())
            m = max(s0.m, d1.m) * max(1.0, d.m)
            d = _V(m, max(s0.e, d1.e) * max(1.0, d.m) + amp(s0.m + d1.m) * d.e + 4 * U * m)
            s0 = vs_pop(vs, a)
        elif name == 'WRAP_PRE':
            et = TRIG_ABS / (2 * math.pi) + 4 * U + (p.e * amp(1.0 / max(abs(c[0]), 1e-30))) / (2 * math.pi)
            p = _V(max(p.m, SQRT2 * p.m + cm), p.e + 4 * U * (SQRT2 * p.m + cm))
            d = _V(1.0, et)
        elif name == 'WRAP_POST':
            span = abs(c[1] - c[0])
            p = _V(max(p.m, cm), max(p.e, amp(span) * d.e + 2 * U * cm))
        elif name == 'PUSH':
            vs_push(vs, a, s0)
            s0 = d.copy()
        elif name == 'SAVEP':
            ps[a] = p.copy()
        elif name == 'LOADP':
            p = ps.get(a, p).copy()
        else:                               # an opcode without a rule must not pass silently
            return INF
        if not (math.isfinite(d.e) and math.isfinite(p.e)):
            return INF
    return d.e


def vs_push(vs, a, s0):
    """PUSH / ELONGATE / EXTRUDE_PRE with operand a: a = 0 leaves the spill levels alone, a = s + 1 spills S0 to VS[s]."""
    if a > 0:
        vs[a - 1] = s0.copy()


def vs_pop(vs, a):
    """A combining op with operand a reloads S0 <- VS[a - 1] (a = 0: the stack below is empty)."""
    return vs[a - 1].copy() if a > 0 and (a - 1) in vs else _V()


def gain(program, coord_max):
    """>= 1: how much more float32 error this tree can accumulate than a tree of the same shape whose constants do
    not amplify.  inf when some node has no finite slope (every sample then goes to float64)."""
    full = error_bound(program, coord_max)
    if not math.isfinite(full):
        return INF
    base = error_bound(program, coord_max, unit_gain=True)
    return max(1.0, full / base) if base > 0 else 1.0

"""Easing functions (the ``e=`` argument of bend_linear / bend_radial / transition_* / wrap_around /
extrude_to), same 34 names as the reference's sdf/ease.py:3-162.

On the GPU an easing is the ``EASE`` opcode with the id below (csrc/sdf_ops.cuh: ease_eval);
the numpy bodies here only serve host-side callers (user closures, plotting).  Every function
carries ``ease_id``; a callable without it cannot be lowered.
"""
import numpy as np

_NAMES = ['linear',
          'in_quad', 'out_quad', 'in_out_quad', 'in_cubic', 'out_cubic', 'in_out_cubic',
          'in_quart', 'out_quart', 'in_out_quart', 'in_quint', 'out_quint', 'in_out_quint',
          'in_sine', 'out_sine', 'in_out_sine', 'in_expo', 'out_expo', 'in_out_expo',
          'in_circ', 'out_circ', 'in_out_circ', 'in_elastic', 'out_elastic', 'in_out_elastic',
          'in_back', 'out_back', 'in_out_back', 'in_bounce', 'out_bounce', 'in_out_bounce',
          'in_square', 'out_square', 'in_out_square']
EASE_ID = {name: i for i, name in enumerate(_NAMES)}


def _power(n):
    def fin(t):
        return t ** n

    def fout(t):
        u = t - 1
        return u ** n + 1 if n % 2 else -(u ** n - 1)

    def finout(t):
        u = t * 2
        v = u - 2
        a = 0.5 * u ** n
        b = 0.5 * (v ** n + 2) if n % 2 else -0.5 * (v ** n - 2)
        return np.where(u < 1, a, b)
    return fin, fout, finout


def linear(t):
    return t


in_quad, _oq, _ioq = _power(2)
in_cubic, out_cubic, in_out_cubic = _power(3)
in_quart, out_quart, in_out_quart = _power(4)
in_quint, out_quint, in_out_quint = _power(5)


def out_quad(t):
    return -t * (t - 2)


def in_out_quad(t):
    u = 2 * t - 1
    return np.where(t < 0.5, 2 * t * t, -0.5 * (u * (u - 2) - 1))


def in_sine(t):
    return 1 - np.cos(t * np.pi / 2)


def out_sine(t):
    return np.sin(t * np.pi / 2)


def in_out_sine(t):
    return -0.5 * (np.cos(np.pi * t) - 1)


def in_expo(t):
    t = np.asarray(t, dtype=float)
    return np.where(t == 0, 0.0, 2 ** (10 * (t - 1)))


def out_expo(t):
    t = np.asarray(t, dtype=float)
    return np.where(t == 1, 1.0, 1 - 2 ** (-10 * t))


def in_out_expo(t):
    t = np.asarray(t, dtype=float)
    lo = 0.5 * 2 ** (20 * t - 10)
    hi = 1 - 0.5 * 2 ** (-20 * t + 10)
    return np.where(t == 0, 0.0, np.where(t == 1, 1.0, np.where(t < 0.5, lo, hi)))


def in_circ(t):
    return 1 - np.sqrt(1 - t * t)


def out_circ(t):
    u = t - 1
    return np.sqrt(1 - u * u)


def in_out_circ(t):
    u = t * 2
    v = u - 2
    with np.errstate(invalid='ignore'):
        return np.where(u < 1, -0.5 * (np.sqrt(1 - u * u) - 1), 0.5 * (np.sqrt(1 - v * v) + 1))


def in_elastic(t, k=0.5):
    u = t - 1
    return -(2 ** (10 * u) * np.sin((u - k / 4) * (2 * np.pi) / k))


def out_elastic(t, k=0.5):
    return 2 ** (-10 * t) * np.sin((t - k / 4) * (2 * np.pi / k)) + 1


def in_out_elastic(t, k=0.5):
    u = t * 2
    v = u - 1
    s = np.sin((v - k / 4) * 2 * np.pi / k)
    return np.where(u < 1, -0.5 * (2 ** (10 * v) * s), 2 ** (-10 * v) * s * 0.5 + 1)


_BACK = 1.70158


def in_back(t):
    return t * t * ((_BACK + 1) * t - _BACK)


def out_back(t):
    u = t - 1
    return u * u * ((_BACK + 1) * u + _BACK) + 1


def in_out_back(t):
    k = _BACK * 1.525
    u = t * 2
    v = u - 2
    return np.where(u < 1, 0.5 * (u * u * ((k + 1) * u - k)), 0.5 * (v * v * ((k + 1) * v + k) + 2))


def out_bounce(t):
    segs = [(4 / 11, (121 / 16, 0.0, 0.0)), (8 / 11, (363 / 40, -99 / 10, 17 / 5)),
            (9 / 10, (4356 / 361, -35442 / 1805, 16061 / 1805)), (np.inf, (54 / 5, -513 / 25, 268 / 25))]
    t = np.asarray(t, dtype=float)
    out = np.zeros_like(t)
    done = np.zeros(t.shape, bool)
    for limit, (a, b, c) in segs:
        m = (~done) & (t < limit)
        out = np.where(m, a * t * t + b * t + c, out)
        done |= m
    return out


def in_bounce(t):
    return 1 - out_bounce(1 - np.asarray(t, dtype=float))


def in_out_bounce(t):
    t = np.asarray(t, dtype=float)
    return np.where(t < 0.5, in_bounce(2 * t) * 0.5, out_bounce(2 * t - 1) * 0.5 + 0.5)


def in_square(t):
    return np.where(np.asarray(t) < 1, 0.0, 1.0)


def out_square(t):
    return np.where(np.asarray(t) > 0, 1.0, 0.0)


def in_out_square(t):
    return np.where(np.asarray(t) < 0.5, 0.0, 1.0)


for _name, _i in EASE_ID.items():
    globals()[_name].ease_id = _i
    globals()[_name].__name__ = _name


def ease_id_of(e):
    """Opcode operand of an easing callable; raises for callables the GPU has no body for."""
    from .ir import NotLowerableError
    i = getattr(e, 'ease_id', None)
    if i is None:
        raise NotLowerableError('easing %r is not one of sdf.ease.* and cannot be lowered to the CUDA evaluator' % (e,))
    return i

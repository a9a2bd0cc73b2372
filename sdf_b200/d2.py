"""2-D SDF object model and node library (host side), mirroring sdf/d2.py:18-298.

A 2-D node evaluates on the (x, y) of the machine's point register and never touches z, so the
same evaluator runs 2-D trees; ``extrude`` / ``extrude_to`` / ``revolve`` (d2.py:261-286) take a
2-D tree back to 3-D.  ``SDF2.__call__`` pads the (N,2) points with z = 0.
"""
import functools

import numpy as np

from . import dn, ease
from . import ir
from .ir import Node

ORIGIN = np.array((0, 0))

X = np.array((1, 0))
Y = np.array((0, 1))

UP = Y

_ops = {}


class SDF2:
    def __init__(self, f=None, node=None):
        self.f = f
        self._node = node if node is not None else Node('opaque', 'OPAQUE', label=getattr(f, '__qualname__', repr(f)))

    def __call__(self, p):
        if self._node.kind == 'opaque':
            return self.f(p).reshape((-1, 1))
        from . import core
        p = np.asarray(p, dtype=np.float64).reshape(-1, 2)
        p3 = np.zeros((len(p), 3))
        p3[:, :2] = p
        return core.evaluate(self, p3).reshape((-1, 1))

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        if name in _ops:
            return functools.partial(_ops[name], self)
        f = self.__dict__.get('f')
        if f is not None:
            return getattr(f, name)
        raise AttributeError(name)

    def __or__(self, other):
        return union(self, other)

    def __and__(self, other):
        return intersection(self, other)

    def __sub__(self, other):
        return difference(self, other)

    def k(self, k=None):
        self._k = k
        return self

    def __setattr__(self, name, value):
        # `_k` is the one attribute a tree can change after it has been built (d3.py:38-40 / dn.py:12): every
        # change -- through k() or by assignment -- invalidates the lowered programs cached on the trees
        if name == '_k':
            ir.MUTATIONS[0] += 1
        object.__setattr__(self, name, value)


def _wrap(result, cls):
    if isinstance(result, Node):
        return cls(node=result)
    if isinstance(result, cls):
        return result
    return cls(f=result)


def sdf2(f):
    @functools.wraps(f)
    def wrapper(*args, **kwargs):
        return _wrap(f(*args, **kwargs), SDF2)
    return wrapper


def op2(f):
    @functools.wraps(f)
    def wrapper(*args, **kwargs):
        return _wrap(f(*args, **kwargs), SDF2)
    _ops[f.__name__] = wrapper
    return wrapper


def op23(f):
    from . import d3

    @functools.wraps(f)
    def wrapper(*args, **kwargs):
        return _wrap(f(*args, **kwargs), d3.SDF3)
    _ops[f.__name__] = wrapper
    return wrapper


def _vec2(v):
    a = np.empty(2, dtype=np.float64)
    a[...] = v
    return a


def _normalize(a):
    a = np.asarray(a, dtype=np.float64)
    return a / np.linalg.norm(a)


# Primitives

@sdf2
def circle(radius=1, center=ORIGIN):
    return Node('prim', 'CIRCLE', [*_vec2(center), radius])


@sdf2
def line(normal=UP, point=ORIGIN):
    return Node('prim', 'LINE2', [*_normalize(normal), *_vec2(point)])


@sdf2
def slab(x0=None, y0=None, x1=None, y1=None, k=None):
    fs = []
    for value, normal, point in ((x0, X, (x0, 0)), (x1, -X, (x1, 0)), (y0, Y, (0, y0)), (y1, -Y, (0, y1))):
        if value is not None:
            fs.append(line(normal, point))
    return intersection(*fs, k=k)


@sdf2
def rectangle(size=1, center=ORIGIN, a=None, b=None):
    if a is not None and b is not None:
        a = np.array(a)
        b = np.array(b)
        size = b - a
        center = a + size / 2
        return rectangle(size, center)
    size = np.array(size)
    return Node('prim', 'RECTANGLE', [*_vec2(center), *_vec2(size / 2)])


@sdf2
def rounded_rectangle(size, radius, center=ORIGIN):
    # (the reference ignores `center` in the closure, d2.py:122-134, and so does this)
    try:
        r0, r1, r2, r3 = radius
    except TypeError:
        r0 = r1 = r2 = r3 = radius
    return Node('prim', 'ROUNDED_RECTANGLE', [*_vec2(np.array(size) / 2), r0, r1, r2, r3])


@sdf2
def equilateral_triangle():
    return Node('prim', 'EQUILATERAL_TRIANGLE', [])


@sdf2
def hexagon(r):
    r *= 3 ** 0.5 / 2
    return Node('prim', 'HEXAGON', [r])


@sdf2
def rounded_x(w, r):
    return Node('prim', 'ROUNDED_X', [w, r])


@sdf2
def polygon(points):
    pts = np.array([np.array(p, dtype=np.float64) for p in points]).reshape(-1, 2)
    return Node('polygon', 'POLYGON', pts.reshape(-1))


@sdf2
def vesica(r, d):
    return Node('prim', 'VESICA', [r, d])


# Positioning (the 3-D opcodes with a neutral third component)

@op2
def translate(other, offset):
    ox, oy = _vec2(offset)
    return Node('transform', 'TRANSLATE', [ox, oy, 0.0], [other])


@op2
def scale(other, factor):
    try:
        x, y = factor
    except TypeError:
        x = y = factor
    m = min(x, y)
    return Node('scale', 'SCALE_P', [x, y, 1.0, m], [other])


@op2
def rotate(other, angle):
    s = np.sin(angle)
    c = np.cos(angle)
    matrix = np.array([
        [c, -s],
        [s, c],
    ]).T
    m3 = np.eye(3)
    m3[:2, :2] = matrix
    return Node('transform', 'ROTATE', m3.reshape(-1), [other])


@op2
def circular_array(other, count):
    angles = [i / count * 2 * np.pi for i in range(count)]
    return union(*[other.rotate(a) for a in angles])


# Alterations

@op2
def elongate(other, size):
    sx, sy = _vec2(size)
    return Node('elongate', 'ELONGATE', [sx, sy, np.inf], [other])      # |z| - inf never wins the max


# 2D => 3D

@op23
def extrude(other, h):
    return Node('extrude', 'EXTRUDE', [h], [other])


@op23
def extrude_to(a, b, h, e=ease.linear):
    return Node('transition', 'EXTRUDE_TO', children=[a, b],
                extra=(('TPARAM_Z', [h]), ease.ease_id_of(e), 'MIX_POP2', [h]))


@op23
def revolve(other, offset=0):
    return Node('transform', 'REVOLVE', [offset], [other])


# Common

union = op2(dn.union)
difference = op2(dn.difference)
intersection = op2(dn.intersection)
blend = op2(dn.blend)
negate = op2(dn.negate)
dilate = op2(dn.dilate)
erode = op2(dn.erode)
shell = op2(dn.shell)


@op2
def repeat(other, spacing, count=None, padding=0):
    return dn.repeat(other, spacing, count, padding, dim=2)

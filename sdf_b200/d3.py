"""3-D SDF object model and node library (host side).

Mirrors the public surface of the reference's ``sdf/d3.py``: the ``SDF3``
callable with ``| & -``, ``.k()``, method chaining through the ``_ops``
registry, the ``sdf3`` / ``op3`` / ``op32`` decorators (d3.py:19-63) and every
factory name in d3.py:92-532.  The difference is what a factory returns: an IR
:class:`~sdf_b200.ir.Node` (constants pre-computed in float64 with the same
numpy expressions the reference uses) instead of a numpy closure.  Evaluation
-- ``SDF3.__call__`` as well as ``generate()`` -- runs on the GPU through the
C-ABI (``sdf_b200._abi``); there is no CPU path.

User code may still decorate its own closure factories with ``@sdf3`` /
``@op3`` (README "writing your own SDFs"): such nodes stay *callable* (the
closure runs in numpy, children evaluate on the GPU) but cannot be meshed.
"""
import functools

import numpy as np

from . import dn, ease
from . import ir
from .ir import Node

# Constants (d3.py:9-17)
ORIGIN = np.array((0, 0, 0))

X = np.array((1, 0, 0))
Y = np.array((0, 1, 0))
Z = np.array((0, 0, 1))

UP = Z

_ops = {}


class SDF3:
    """Callable signed-distance function of 3-D points (d3.py:21-46)."""

    def __init__(self, f=None, node=None):
        self.f = f
        self._node = node if node is not None else Node('opaque', 'OPAQUE', label=getattr(f, '__qualname__', repr(f)))

    def __call__(self, p):
        if self._node.kind == 'opaque':
            return self.f(p).reshape((-1, 1))
        from . import core
        return core.evaluate(self, p).reshape((-1, 1))

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

    def generate(self, *args, **kwargs):
        from . import core
        return core.generate(self, *args, **kwargs)

    def save(self, path, *args, **kwargs):
        from . import core
        return core.save(path, self, *args, **kwargs)

    def show_slice(self, *args, **kwargs):
        from . import core
        return core.show_slice(self, *args, **kwargs)


def _wrap(result, cls):
    if isinstance(result, Node):
        return cls(node=result)
    if isinstance(result, cls):
        return result       # slab() -> intersection(...), rotate_to() -> rotate(...): reuse the inner node
    return cls(f=result)    # user closure: callable, not lowerable


def sdf3(f):
    @functools.wraps(f)
    def wrapper(*args, **kwargs):
        return _wrap(f(*args, **kwargs), SDF3)
    return wrapper


def op3(f):
    @functools.wraps(f)
    def wrapper(*args, **kwargs):
        return _wrap(f(*args, **kwargs), SDF3)
    _ops[f.__name__] = wrapper
    return wrapper


def op32(f):
    from . import d2

    @functools.wraps(f)
    def wrapper(*args, **kwargs):
        return _wrap(f(*args, **kwargs), d2.SDF2)
    _ops[f.__name__] = wrapper
    return wrapper


# Helpers (host-side constant preparation only)

def _normalize(a):
    a = np.asarray(a, dtype=np.float64)
    return a / np.linalg.norm(a)


def _perpendicular(v):
    if v[1] == 0 and v[2] == 0:
        if v[0] == 0:
            raise ValueError('zero vector')
        return np.cross(v, [0, 1, 0])
    return np.cross(v, [1, 0, 0])


def _vec3(v):
    a = np.empty(3, dtype=np.float64)
    a[...] = v          # broadcasts scalars, as numpy does inside the reference closures
    return a


# Primitives

@sdf3
def sphere(radius=1, center=ORIGIN):
    return Node('prim', 'SPHERE', [*_vec3(center), radius])


@sdf3
def plane(normal=UP, point=ORIGIN):
    return Node('prim', 'PLANE', [*_normalize(normal), *_vec3(point)])


@sdf3
def slab(x0=None, y0=None, z0=None, x1=None, y1=None, z1=None, k=None):
    fs = []
    if x0 is not None:
        fs.append(plane(X, (x0, 0, 0)))
    if x1 is not None:
        fs.append(plane(-X, (x1, 0, 0)))
    if y0 is not None:
        fs.append(plane(Y, (0, y0, 0)))
    if y1 is not None:
        fs.append(plane(-Y, (0, y1, 0)))
    if z0 is not None:
        fs.append(plane(Z, (0, 0, z0)))
    if z1 is not None:
        fs.append(plane(-Z, (0, 0, z1)))
    return intersection(*fs, k=k)


@sdf3
def box(size=1, center=ORIGIN, a=None, b=None):
    if a is not None and b is not None:
        a = np.array(a)
        b = np.array(b)
        size = b - a
        center = a + size / 2
        return box(size, center)
    size = np.array(size)
    return Node('prim', 'BOX', [*_vec3(center), *_vec3(size / 2)])


@sdf3
def rounded_box(size, radius):
    size = np.array(size)
    return Node('prim', 'ROUNDED_BOX', [*_vec3(size / 2), radius])


@sdf3
def wireframe_box(size, thickness):
    size = np.array(size)
    return Node('prim', 'WIREFRAME_BOX', [*_vec3(size / 2), thickness])


@sdf3
def torus(r1, r2):
    return Node('prim', 'TORUS', [r1, r2])


@sdf3
def capsule(a, b, radius):
    return Node('prim', 'CAPSULE', [*_vec3(np.array(a)), *_vec3(np.array(b)), radius])


@sdf3
def cylinder(radius):
    return Node('prim', 'CYLINDER', [radius])


@sdf3
def capped_cylinder(a, b, radius):
    return Node('prim', 'CAPPED_CYLINDER', [*_vec3(np.array(a)), *_vec3(np.array(b)), radius])


@sdf3
def rounded_cylinder(ra, rb, h):
    return Node('prim', 'ROUNDED_CYLINDER', [ra, rb, h])


@sdf3
def capped_cone(a, b, ra, rb):
    return Node('prim', 'CAPPED_CONE', [*_vec3(np.array(a)), *_vec3(np.array(b)), ra, rb])


@sdf3
def rounded_cone(r1, r2, h):
    return Node('prim', 'ROUNDED_CONE', [r1, r2, h])


@sdf3
def ellipsoid(size):
    return Node('prim', 'ELLIPSOID', _vec3(np.array(size)))


@sdf3
def pyramid(h):
    return Node('prim', 'PYRAMID', [h])


# Platonic solids

@sdf3
def tetrahedron(r):
    return Node('prim', 'TETRAHEDRON', [r])


@sdf3
def octahedron(r):
    return Node('prim', 'OCTAHEDRON', [r, np.tan(np.radians(30))])


@sdf3
def dodecahedron(r):
    x, y, z = _normalize(((1 + np.sqrt(5)) / 2, 1, 0))
    return Node('prim', 'DODECAHEDRON', [r, x, y, z])


@sdf3
def icosahedron(r):
    r *= 0.8506507174597755
    x, y, z = _normalize(((np.sqrt(5) + 3) / 2, 1, 0))
    w = np.sqrt(3) / 3
    return Node('prim', 'ICOSAHEDRON', [r, x, y, z, w])


# Positioning

@op3
def translate(other, offset):
    return Node('transform', 'TRANSLATE', _vec3(offset), [other])


@op3
def scale(other, factor):
    try:
        x, y, z = factor
    except TypeError:
        x = y = z = factor
    m = min(x, min(y, z))
    return Node('scale', 'SCALE_P', [x, y, z, m], [other])


@op3
def rotate(other, angle, vector=Z):
    x, y, z = _normalize(vector)
    s = np.sin(angle)
    c = np.cos(angle)
    m = 1 - c
    matrix = np.array([
        [m*x*x + c, m*x*y + z*s, m*z*x - y*s],
        [m*x*y - z*s, m*y*y + c, m*y*z + x*s],
        [m*z*x + y*s, m*y*z - x*s, m*z*z + c],
    ]).T
    return Node('transform', 'ROTATE', matrix.reshape(-1), [other])


@op3
def rotate_to(other, a, b):
    a = _normalize(np.array(a))
    b = _normalize(np.array(b))
    dot = np.dot(b, a)
    if dot == 1:
        return other
    if dot == -1:
        return rotate(other, np.pi, _perpendicular(a))
    angle = np.arccos(dot)
    v = _normalize(np.cross(b, a))
    return rotate(other, angle, v)


@op3
def orient(other, axis):
    return rotate_to(other, UP, axis)


@op3
def circular_array(other, count, offset=0):
    da = 2 * np.pi / count
    return Node('circular', 'CIRCULAR_PRE', [da, offset], [other])


# Alterations

@op3
def elongate(other, size):
    return Node('elongate', 'ELONGATE', _vec3(size), [other])


@op3
def twist(other, k):
    return Node('transform', 'TWIST', [k], [other])


@op3
def bend(other, k):
    return Node('transform', 'BEND', [k], [other])


@op3
def bend_linear(other, p0, p1, v, e=ease.linear):
    p0 = np.array(p0)
    p1 = np.array(p1)
    v = -np.array(v)
    ab = p1 - p0
    steps = [('TPARAM_LINEAR', 0, [*_vec3(p0), *_vec3(ab)]), ('EASE', ease.ease_id_of(e), []), ('DISPLACE', 0, _vec3(v))]
    return Node('warp', 'BEND_LINEAR', children=[other], extra=steps)


@op3
def bend_radial(other, r0, r1, dz, e=ease.linear):
    steps = [('TPARAM_RADIAL', 0, [r0, r1]), ('EASE', ease.ease_id_of(e), []), ('DISPLACE', 0, [0.0, 0.0, -dz])]
    return Node('warp', 'BEND_RADIAL', children=[other], extra=steps)


@op3
def transition_linear(f0, f1, p0=-Z, p1=Z, e=ease.linear):
    p0 = np.array(p0)
    p1 = np.array(p1)
    ab = p1 - p0
    return Node('transition', 'TRANSITION_LINEAR', children=[f0, f1],
                extra=(('TPARAM_LINEAR', [*_vec3(p0), *_vec3(ab)]), ease.ease_id_of(e), 'LERP_POP2', None))


@op3
def transition_radial(f0, f1, r0=0, r1=1, e=ease.linear):
    return Node('transition', 'TRANSITION_RADIAL', children=[f0, f1],
                extra=(('TPARAM_RADIAL', [r0, r1]), ease.ease_id_of(e), 'LERP_POP2', None))


@op3
def wrap_around(other, x0, x1, r=None, e=ease.linear):
    p0 = X * x0
    p1 = X * x1
    if r is None:
        r = np.linalg.norm(p1 - p0) / (2 * np.pi)
    steps = [('WRAP_PRE', 0, [r]), ('EASE', ease.ease_id_of(e), []), ('WRAP_POST', 0, [x0, x1])]
    return Node('warp', 'WRAP_AROUND', children=[other], extra=steps)


# 3D => 2D

@op32
def slice(other):
    s = slab(z0=-1e-9, z1=1e-9)
    a = other & s
    b = other.negate() & s
    return Node('slice', 'SLICE', children=[a, b])


# Common (dn.py)

union = op3(dn.union)
difference = op3(dn.difference)
intersection = op3(dn.intersection)
blend = op3(dn.blend)
negate = op3(dn.negate)
dilate = op3(dn.dilate)
erode = op3(dn.erode)
shell = op3(dn.shell)


@op3
def repeat(other, spacing, count=None, padding=0):
    return dn.repeat(other, spacing, count, padding, dim=3)

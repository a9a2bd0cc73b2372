"""2-D SDF object model (host side), mirroring sdf/d2.py:18-55.

Only the object protocol, the decorators and the registry exist in this round;
the 2-D primitive opcodes and extrude/revolve are SURVEY.md section 8(f) rank 1
("next").  Factories that are not lowered yet raise NotImplementedError on use,
they never fall back to a CPU path.
"""
import functools

import numpy as np

from . import dn
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
        raise NotImplementedError('2-D evaluation is not lowered to the CUDA evaluator yet')

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

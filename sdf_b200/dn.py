"""Dimension-agnostic combinators (host side): IR records for sdf/dn.py:7-114.

Each function returns a :class:`~sdf_b200.ir.Node`; the 2-D / 3-D modules wrap
them with ``op2`` / ``op3`` exactly as the reference does (d3.py:524-532).
The smooth-k choice ``K = k or getattr(b, '_k', None)`` (dn.py:12,27,42,57) is
taken at lowering time so that a ``.k()`` applied after the fold was built is
honoured, as in the reference where it is read at call time.
"""
import itertools

import numpy as np

from .ir import Node


def union(a, *bs, k=None):
    return Node('fold', 'union', children=[a, *bs], extra=k)


def difference(a, *bs, k=None):
    return Node('fold', 'difference', children=[a, *bs], extra=k)


def intersection(a, *bs, k=None):
    return Node('fold', 'intersection', children=[a, *bs], extra=k)


def blend(a, *bs, k=0.5):
    return Node('blend', 'blend', children=[a, *bs], extra=k)


def negate(other):
    return Node('value', 'NEGATE', children=[other])


def dilate(other, r):
    return Node('value', 'ADDC', [-r], [other], label='dilate')


def erode(other, r):
    return Node('value', 'ADDC', [r], [other], label='erode')


def shell(other, thickness):
    return Node('value', 'SHELL', [thickness], [other])


def repeat(other, spacing, count=None, padding=0, dim=3):
    """dn.py:82-114.  ``dim`` is the point dimension the reference reads from
    ``p.shape[-1]`` at call time."""
    has_count = count is not None
    cnt = np.zeros(3)
    if has_count:
        cnt[:dim] = np.array(count, dtype=np.float64)
    sp = np.zeros(3)
    sp[:dim] = np.array(spacing, dtype=np.float64)
    try:
        pad = [padding[i] for i in range(dim)]
    except Exception:
        pad = [padding] * dim
    for i in range(dim):
        if sp[i] == 0:
            pad[i] = 0
    pad = list(pad) + [0] * (3 - dim)
    neighbours = list(itertools.product(*[range(-p, p + 1) for p in pad]))
    return Node('repeat', 'REPEAT_CELL', children=[other],
                extra=(list(sp), list(cnt), 1.0 if has_count else 0.0, neighbours))

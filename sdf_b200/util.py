"""Names that user scripts pick up through ``from sdf import *`` (reference sdf/util.py:1-6)."""
from math import degrees, pi, radians  # noqa: F401

__all__ = ['pi', 'degrees', 'radians']

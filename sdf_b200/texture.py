"""Names the reference star-exports for its texture nodes (sdf/__init__.py:11-18: ``Mesh`` from
sdf/mesh.py, ``text / image / measure_text / measure_image`` from sdf/text.py).

They sample a PIL image or an OpenVDB level set inside the closure -- data the GPU evaluator has no
opcode for -- and are outside the mesh-generation hot path this package accelerates (DESIGN.md
section 7).  The names exist so that ``from sdf import *`` scripts import cleanly; using one raises.
"""


def _out_of_scope(name):
    raise NotImplementedError(
        '%s() is not provided by sdf_b200: texture-backed nodes (PIL / OpenVDB) are outside the GPU '
        'mesh-generation path; build the tree from the analytic primitives and operators' % name)


def measure_text(name, text, width=None, height=None):
    _out_of_scope('measure_text')


def measure_image(thing, width=None, height=None):
    _out_of_scope('measure_image')


def text(font_name, text, width=None, height=None, pixels=None, points=512):
    _out_of_scope('text')


def image(thing, width=None, height=None, pixels=None):
    _out_of_scope('image')


class Mesh:
    """sdf/mesh.py:9 (a triangle mesh turned into an SDF through pyopenvdb)."""

    def __init__(self, *args, **kwargs):
        _out_of_scope('Mesh')

    @staticmethod
    def from_file(*args, **kwargs):
        _out_of_scope('Mesh.from_file')

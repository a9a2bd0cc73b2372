"""Expression-tree IR and its lowering to the evaluator's linear program.

The reference builds a tree of numpy closures (sdf/d3.py:48-63 wrap every
factory's closure in ``SDF3``); evaluation is a recursive Python call chain
(sdf/d3.py:24-25).  Here every factory records a :class:`Node` instead, and
:func:`lower` flattens the tree into ``(uint32 code[], float64 consts[])`` for
the CUDA interpreter (csrc/sdf_interp.cuh) behind the C-ABI
``sdfb200_program_create`` (include/sdfb200.h).

Machine model (per sample point): registers ``P`` (x,y,z), ``D`` (distance of
the sub-tree just evaluated), ``S0`` (cached top of the value stack), ``R0``
(cached point slot 0); a value stack ``VS[]`` and point slots ``PS[]`` spill to
per-thread local memory.  One instruction word = ``op | a << 8 | ci << 16``
where ``ci`` indexes ``consts``.
"""
import numpy as np

# --------------------------------------------------------------------------
# opcode table: name -> (code, number of constants).  The C header
# csrc/sdf_opcodes.h is generated from this table (tools/gen_opcodes.py).

_OPS = [
    # primitives: read P, write D                       constants
    ('SPHERE', 4),            # cx cy cz r                         d3.py:92-96
    ('PLANE', 6),             # nx ny nz px py pz                  d3.py:98-103
    ('BOX', 6),               # cx cy cz hx hy hz                  d3.py:122-134
    ('ROUNDED_BOX', 4),       # hx hy hz r                         d3.py:136-142
    ('WIREFRAME_BOX', 4),     # hx hy hz t                         d3.py:144-155
    ('TORUS', 2),             # r1 r2                              d3.py:157-165
    ('CAPSULE', 7),           # ax ay az bx by bz r                d3.py:167-176
    ('CYLINDER', 1),          # r                                  d3.py:178-182
    ('CAPPED_CYLINDER', 7),   # ax ay az bx by bz r                d3.py:184-204
    ('ROUNDED_CYLINDER', 3),  # ra rb h                            d3.py:206-215
    ('CAPPED_CONE', 8),       # ax ay az bx by bz ra rb            d3.py:217-237
    ('ROUNDED_CONE', 3),      # r1 r2 h                            d3.py:239-250
    ('ELLIPSOID', 3),         # sx sy sz                           d3.py:252-259
    ('PYRAMID', 1),           # h                                  d3.py:261-282
    ('TETRAHEDRON', 1),       # r                                  d3.py:286-293
    ('OCTAHEDRON', 2),        # r tan(radians(30))                 d3.py:295-299
    ('DODECAHEDRON', 4),      # r x y z                            d3.py:301-311
    ('ICOSAHEDRON', 5),       # r' x y z w                         d3.py:313-325
    # point transforms: P <- T(P)
    ('TRANSLATE', 3),         # ox oy oz                           d3.py:329-333
    ('SCALE_P', 3),           # sx sy sz  (P / s)                  d3.py:335-345
    ('ROTATE', 9),            # row-major M, P' = P @ M            d3.py:347-360
    ('TWIST', 1),             # k                                  d3.py:407-419
    ('BEND', 1),              # k                                  d3.py:421-433
    ('ELONGATE', 3),          # sx sy sz ; pushes w on the value stack  d3.py:396-405
    ('CIRCULAR_PRE', 1),      # da ; P <- first copy, PS[a] <- second   d3.py:379-392
    ('REPEAT_CELL', 12),      # spacing(3) count(3) has_count(3 flags -> 1) nb(3) ...  dn.py:82-114
    # value ops: D <- g(D)
    ('MULC', 1),              # m                                  d3.py:344
    ('NEGATE', 0),            #                                    dn.py:62-65
    ('ADDC', 1),              # c (dilate: -r, erode: +r)          dn.py:67-75
    ('SHELL', 1),             # thickness                          dn.py:77-80
    ('ADD_POP', 0),           # D <- pop + D (elongate tail)       d3.py:405
    # CSG folds: D <- op(S0, D), pops the value stack
    ('UNION', 0),             #                                    dn.py:14
    ('UNION_K', 1),           # K                                  dn.py:16-18
    ('DIFFERENCE', 0),        #                                    dn.py:29
    ('DIFFERENCE_K', 1),      # K                                  dn.py:31-33
    ('INTERSECTION', 0),      #                                    dn.py:44
    ('INTERSECTION_K', 1),    # K                                  dn.py:46-48
    ('BLEND', 1),             # K                                  dn.py:52-60
    # stack / slots
    ('PUSH', 0),              # a = 0: S0 <- D ; a = s+1: VS[s] <- S0, S0 <- D
    ('SAVEP', 0),             # a = slot (0 = register slot)
    ('LOADP', 0),
]

OPCODE = {name: i + 1 for i, (name, _) in enumerate(_OPS)}
NCONST = {name: n for name, n in _OPS}
OPNAME = {v: k for k, v in OPCODE.items()}

FOLD_OPS = {'union': ('UNION', 'UNION_K'), 'difference': ('DIFFERENCE', 'DIFFERENCE_K'),
            'intersection': ('INTERSECTION', 'INTERSECTION_K')}

MAX_CODE = 512       # instruction words that fit the kernel-parameter program image
MAX_CONSTS = 1024
MAX_VSTACK = 16
MAX_PSLOTS = 8


class NotLowerableError(TypeError):
    """Raised when a tree contains a node the CUDA evaluator has no opcode for.

    ``generate()`` has no CPU fallback: an opaque ``@sdf3`` closure (README
    "writing your own SDFs") can still be *called* on points, but cannot be
    meshed on the GPU.
    """


class Node:
    """One vertex of the expression tree.

    kind: 'prim' | 'transform' | 'value' | 'fold' | 'blend' | 'circular' |
          'elongate' | 'scale' | 'repeat' | 'opaque'
    op:   opcode name (or the fold name for kind == 'fold')
    consts: flat list of python floats (float64)
    children: list of SDF objects (objects, not Nodes, so that ``.k()`` set
          after construction is still honoured, as in dn.py:12)
    """
    __slots__ = ('kind', 'op', 'consts', 'children', 'extra', 'label')

    def __init__(self, kind, op, consts=(), children=(), extra=None, label=None):
        self.kind = kind
        self.op = op
        self.consts = [float(c) for c in np.asarray(consts, dtype=np.float64).reshape(-1)]
        self.children = list(children)
        self.extra = extra
        self.label = label or op.lower()


class Program:
    """Lowered program: what crosses the C-ABI."""

    def __init__(self, code, consts, vdepth, pslots, scale_hint):
        self.code = np.ascontiguousarray(code, dtype=np.uint32)
        self.consts = np.ascontiguousarray(consts, dtype=np.float64)
        self.vdepth = int(vdepth)
        self.pslots = int(pslots)
        self.scale_hint = float(scale_hint)

    def disassemble(self):
        out = []
        for w in self.code:
            op, a, ci = int(w) & 0xff, (int(w) >> 8) & 0xff, int(w) >> 16
            name = OPNAME[op]
            n = NCONST[name]
            out.append('%-16s a=%d  %s' % (name, a, np.array2string(self.consts[ci:ci + n], precision=6)))
        return '\n'.join(out)


def node_of(obj):
    node = getattr(obj, '_node', None)
    if node is None:
        raise NotLowerableError(
            'SDF %r has no IR node: it wraps an opaque Python closure and cannot be lowered to '
            'the CUDA evaluator (no CPU fallback exists for generate())' % (obj,))
    return node


def _clobbers(obj):
    n = node_of(obj)
    if n.kind == 'prim':
        return False
    if n.kind == 'value':
        return _clobbers(n.children[0])
    if n.kind in ('fold', 'blend'):
        return any(_clobbers(c) for c in n.children)
    return True


class _Lowerer:
    def __init__(self):
        self.code = []
        self.consts = []
        self.vdepth = 0
        self.max_vdepth = 0
        self.pslot = 0
        self.max_pslot = 0

    # -- emission helpers
    def emit(self, name, a=0, consts=None):
        if consts is None:
            consts = ()
        assert len(consts) == NCONST[name], (name, len(consts))
        ci = len(self.consts)
        self.consts.extend(float(c) for c in consts)
        if not (0 <= a < 256 and ci < 65536):
            raise NotLowerableError('tree too deep / too many constants for the instruction encoding '
                                    '(operand %d, constant index %d)' % (a, ci))
        self.code.append(OPCODE[name] | (a << 8) | (ci << 16))

    def push(self):
        self.emit('PUSH', a=self.vdepth)     # a = 0: S0 <- D ; a = s+1 spills S0 to VS[s]
        self.vdepth += 1
        self.max_vdepth = max(self.max_vdepth, self.vdepth)

    def combine(self, name, consts=()):
        self.vdepth -= 1
        # a = 0: stack below is empty ; a = s+1: reload S0 <- VS[s]
        self.emit(name, a=self.vdepth, consts=consts)

    def alloc_pslot(self):
        s = self.pslot
        self.pslot += 1
        self.max_pslot = max(self.max_pslot, self.pslot)
        return s

    def free_pslot(self):
        self.pslot -= 1

    # -- tree walk
    def node(self, obj):
        n = node_of(obj)
        kind = n.kind
        if kind == 'opaque':
            raise NotLowerableError('node %r is an opaque Python closure; the CUDA evaluator has '
                                    'no opcode for it' % (n.label,))
        if kind == 'prim':
            self.emit(n.op, consts=n.consts)
        elif kind == 'transform':
            self.emit(n.op, consts=n.consts)
            self.node(n.children[0])
        elif kind == 'value':
            self.node(n.children[0])
            self.emit(n.op, consts=n.consts)
        elif kind == 'scale':
            self.emit('SCALE_P', consts=n.consts[:3])
            self.node(n.children[0])
            self.emit('MULC', consts=n.consts[3:4])
        elif kind == 'elongate':
            self.emit('ELONGATE', a=self.vdepth, consts=n.consts)   # pushes w
            self.vdepth += 1
            self.max_vdepth = max(self.max_vdepth, self.vdepth)
            self.node(n.children[0])
            self.combine('ADD_POP')
        elif kind == 'fold':
            self.fold(n)
        elif kind == 'blend':
            self.fold(n, blend=True)
        elif kind == 'circular':
            self.circular(n)
        elif kind == 'repeat':
            self.repeat(n)
        else:
            raise NotLowerableError('unknown node kind %r' % (kind,))

    def fold(self, n, blend=False):
        kids = n.children
        need_save = any(_clobbers(c) for c in kids[:-1])
        slot = None
        if need_save:
            slot = self.alloc_pslot()
            self.emit('SAVEP', a=slot)
        self.node(kids[0])
        dirty = _clobbers(kids[0])
        k_op = n.extra            # the fold-level k (dn.py:12: ``k or getattr(b, '_k', None)``)
        for b in kids[1:]:
            self.push()
            if dirty:
                self.emit('LOADP', a=slot)
            self.node(b)
            dirty = _clobbers(b)
            K = k_op or getattr(b, '_k', None)
            if blend:
                self.combine('BLEND', consts=(K,))
            elif K is None:
                self.combine(FOLD_OPS[n.op][0])
            else:
                self.combine(FOLD_OPS[n.op][1], consts=(K,))
        if need_save:
            self.free_pslot()

    def circular(self, n):
        # d3.py:379-392: other = other.translate(X * offset); two rotated copies; min
        slot = self.alloc_pslot()
        da, offset = n.consts
        self.emit('CIRCULAR_PRE', a=slot, consts=(da,))
        self.emit('TRANSLATE', consts=(offset, 0.0, 0.0))
        self.node(n.children[0])
        self.push()
        self.emit('LOADP', a=slot)
        self.emit('TRANSLATE', consts=(offset, 0.0, 0.0))
        self.node(n.children[0])
        self.combine('UNION')
        self.free_pslot()

    def repeat(self, n):
        # dn.py:82-114: min over neighbour cells of other(p - spacing * (index + nb))
        slot = self.alloc_pslot()
        self.emit('SAVEP', a=slot)
        spacing, count, has_count, neighbours = n.extra
        first = True
        for nb in neighbours:
            if not first:
                self.push()
                self.emit('LOADP', a=slot)
            consts = list(spacing) + list(count) + [has_count, nb[0], nb[1], nb[2], 0.0, 0.0]
            self.emit('REPEAT_CELL', consts=consts)
            self.node(n.children[0])
            if not first:
                self.combine('UNION')
            first = False
        self.free_pslot()


def lower(obj):
    """Lower the tree rooted at SDF object ``obj`` to a :class:`Program`."""
    lw = _Lowerer()
    lw.node(obj)
    if len(lw.code) > MAX_CODE or len(lw.consts) > MAX_CONSTS:
        raise NotLowerableError('program too large for the kernel-parameter image: %d words / %d '
                                'constants (limits %d / %d)' % (len(lw.code), len(lw.consts),
                                                                MAX_CODE, MAX_CONSTS))
    if lw.max_vdepth > MAX_VSTACK or lw.max_pslot > MAX_PSLOTS:
        raise NotLowerableError('tree too deep: value stack %d (max %d), point slots %d (max %d)'
                                % (lw.max_vdepth, MAX_VSTACK, lw.max_pslot, MAX_PSLOTS))
    consts = np.array(lw.consts, dtype=np.float64)
    finite = consts[np.isfinite(consts)]
    scale = float(np.max(np.abs(finite))) if finite.size else 1.0
    return Program(np.array(lw.code, dtype=np.uint32), consts, lw.max_vdepth, lw.max_pslot, scale)

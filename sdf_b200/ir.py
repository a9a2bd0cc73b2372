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
    # 2-D primitives (read P.xy, write D)                               sdf/d2.py
    ('CIRCLE', 3),            # cx cy r                            d2.py:76-80
    ('LINE2', 4),             # nx ny px py                        d2.py:82-87
    ('RECTANGLE', 4),         # cx cy hx hy                        d2.py:102-114
    ('ROUNDED_RECTANGLE', 6), # hx hy r0 r1 r2 r3                  d2.py:116-135
    ('EQUILATERAL_TRIANGLE', 0),  #                                d2.py:137-153
    ('HEXAGON', 1),           # r * sqrt(3)/2                      d2.py:155-167
    ('ROUNDED_X', 2),         # w r                                d2.py:169-175
    ('POLYGON', -1),          # a = n points, 2n constants         d2.py:177-198
    ('VESICA', 2),            # r d                                d2.py:200-209
    # 2-D <-> 3-D
    ('REVOLVE', 1),           # offset ; P <- (|xy| - offset, z)    d2.py:280-286
    ('EXTRUDE_PRE', 1),       # h ; pushes |z| - h/2               d2.py:261-266
    ('EXTRUDE_POST', 0),      # D <- min(max(D,w),0)+|max((D,w),0)|, pops
    ('SLICE_PRE', 0),         # P.z <- 0                            d3.py:506-520
    ('SLICE_SELECT', 0),      # D <- (A <= 0) ? -D : A with A = pop
    # eased warps / transitions (the `e=` parameter)                    sdf/ease.py
    ('TPARAM_LINEAR', 6),     # p0(3) ab(3) ; D <- clip(dot(p-p0,ab)/dot(ab,ab),0,1)   d3.py:442,464
    ('TPARAM_RADIAL', 2),     # r0 r1 ; D <- clip((|xy|-r0)/(r1-r0),0,1)               d3.py:453-454,476-477
    ('TPARAM_Z', 1),          # h ; D <- clip(z/h,-.5,.5)+.5                          d2.py:273
    ('EASE', 0),              # a = ease id ; D <- e(D)            ease.py:3-162
    ('DISPLACE', 3),          # v ; P <- P + D*v                   d3.py:444,455
    ('LERP_POP2', 0),         # D <- t*d2 + (1-t)*d1, pops 2       d3.py:466,479
    ('MIX_POP2', 0),          # D <- d1 + (d2-d1)*t, pops 2        d2.py:274
    ('WRAP_PRE', 1),          # r ; D <- t, P.y <- -(|xy|-r)       d3.py:490-497
    ('WRAP_POST', 2),         # x0 x1 ; P.x <- x0 + (x1-x0)*D      d3.py:498
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

MUTATIONS = [0]      # bumped whenever some SDF object's `_k` changes (d3.py / d2.py __setattr__): cached lowerings are stale

MAX_CODE = 8192      # instruction words (the kernel-parameter image of the interpreter holds 512 / 1024;
MAX_CONSTS = 32768   #  larger programs run only on the run-time specialised kernels)
IMG_MAX_CODE = 512   # capacity of the interpreter's kernel-parameter program image
IMG_MAX_CONSTS = 1024
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
            n = NCONST[name] if NCONST[name] >= 0 else 2 * a
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
    if n.kind in ('prim', 'polygon'):
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
        assert NCONST[name] < 0 or len(consts) == NCONST[name], (name, len(consts))
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
        elif kind == 'extrude':                     # d2.py:261-266
            self.emit('EXTRUDE_PRE', a=self.vdepth, consts=n.consts)    # pushes |z| - h/2
            self.vdepth += 1
            self.max_vdepth = max(self.max_vdepth, self.vdepth)
            self.node(n.children[0])
            self.combine('EXTRUDE_POST')
        elif kind == 'polygon':
            self.emit('POLYGON', a=len(n.consts) // 2, consts=n.consts)
        elif kind == 'warp':                        # bend_linear / bend_radial / wrap_around
            for op, a, consts in n.extra:
                self.emit(op, a=a, consts=consts)
            self.node(n.children[0])
        elif kind == 'transition':                  # transition_linear/_radial, extrude_to
            self.transition(n)
        elif kind == 'slice':
            self.slice(n)
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

    def fold(self, n, blend=False, shared_slot=None):
        """``shared_slot``: this fold is the FIRST operand of an enclosing fold, evaluated at the point
        that fold has just saved -- it restores from the same slot instead of saving P again (keeps
        the slot depth of ``f |= ...; f -= ...`` chains at one)."""
        kids = n.children
        need_save = any(_clobbers(c) for c in kids[:-1])
        slot = shared_slot
        own = need_save and shared_slot is None
        if own:
            slot = self.alloc_pslot()
            self.emit('SAVEP', a=slot)
        first = node_of(kids[0])
        if need_save and first.kind in ('fold', 'blend'):
            self.fold(first, blend=(first.kind == 'blend'), shared_slot=slot)
        else:
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
        if own:
            self.free_pslot()

    def transition(self, n):
        """d1 = f0(p); d2 = f1(p); t = e(param(p)); combine(d1, d2, t)  (d3.py:458-481, d2.py:268-278)"""
        f0, f1 = n.children
        tparam, ease_id, mix, tail = n.extra
        slot = self.alloc_pslot()
        self.emit('SAVEP', a=slot)
        self.node(f0)
        self.push()
        self.emit('LOADP', a=slot)
        self.node(f1)
        self.push()
        self.emit('LOADP', a=slot)
        self.emit(tparam[0], consts=tparam[1])
        self.emit('EASE', a=ease_id)
        self.vdepth -= 2
        self.emit(mix, a=self.vdepth)
        self.free_pslot()
        if tail is not None:                        # extrude_to finishes like extrude
            self.emit('EXTRUDE_PRE', a=self.vdepth, consts=tail)
            # EXTRUDE_PRE pushes w and leaves D untouched; EXTRUDE_POST combines
            self.vdepth += 1
            self.max_vdepth = max(self.max_vdepth, self.vdepth)
            self.combine('EXTRUDE_POST')

    def slice(self, n):
        """d3.py:506-520: A = (other & s)(x,y,0); B = -(other.negate() & s)(x,y,0); A <= 0 ? B : A"""
        a_tree, b_tree = n.children
        self.emit('SLICE_PRE')
        slot = self.alloc_pslot()
        self.emit('SAVEP', a=slot)
        self.node(a_tree)
        self.push()
        self.emit('LOADP', a=slot)
        self.node(b_tree)
        self.combine('SLICE_SELECT')
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

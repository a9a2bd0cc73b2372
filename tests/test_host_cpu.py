"""CPU suite, part 3: host logic (IR lowering, lattice, batching, slab partition, STL sink) and the
C-ABI library surface.  No compute call is made here (there is no GPU in this container)."""
import ctypes
import io
import os
import re
import struct

import numpy as np
import pytest

import models
import sdf_b200
from sdf_b200 import _abi, core, ir, multi, stl
from oracle import sdf_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ IR / lowering

def test_every_corpus_model_lowers():
    for name, f in models.build(sdf_b200).items():
        prog = ir.lower(f)
        assert 1 <= len(prog.code) <= ir.MAX_CODE, name
        assert prog.vdepth <= ir.MAX_VSTACK and prog.pslots <= ir.MAX_PSLOTS
        for w in prog.code:
            op, ci = int(w) & 0xff, int(w) >> 16
            assert op in ir.OPNAME
            assert ci + ir.NCONST[ir.OPNAME[op]] <= len(prog.consts)
        assert prog.disassemble()


def test_config_programs_are_small():
    sizes = {k: len(ir.lower(f(sdf_b200)).code) for k, f in models.CONFIGS.items()}
    assert sizes['sphere'] == 1 and sizes['knurling'] < 64, sizes


def test_k_selection_rule():
    """dn.py:12: K = k or getattr(b, '_k', None) -- op-level k wins, else the RIGHT operand's .k();
    the left operand's _k is never read; k=0 is falsy; .k() set after the fold is still honoured."""
    s, b = sdf_b200.sphere(1), sdf_b200.box(1)

    def ops(f):
        return [ir.OPNAME[int(w) & 0xff] for w in ir.lower(f).code]
    assert 'UNION' in ops(s | b) and 'UNION_K' not in ops(s | b)
    assert 'UNION_K' in ops(s | sdf_b200.box(1).k(0.2))
    assert 'UNION_K' not in ops(sdf_b200.sphere(1).k(0.2) | b)            # left operand's k ignored
    assert 'DIFFERENCE_K' in ops(sdf_b200.difference(s, b, k=0.3))
    assert 'UNION' in ops(sdf_b200.union(s, b, k=0))                      # k=0 is falsy
    late = sdf_b200.box(1)
    f = s & late
    late.k(0.25)
    prog = ir.lower(f)
    assert 'INTERSECTION_K' in ops(f) and 0.25 in prog.consts


def test_point_save_restore_only_when_needed():
    names = lambda f: [ir.OPNAME[int(w) & 0xff] for w in ir.lower(f).code]
    assert 'SAVEP' not in names(sdf_b200.sphere(1) | sdf_b200.box(1))                # no child moves P
    n = names(sdf_b200.sphere(1).translate((1, 0, 0)) | sdf_b200.box(1))
    assert n.count('SAVEP') == 1 and n.count('LOADP') == 1
    n = names(sdf_b200.sphere(1) | sdf_b200.box(1).translate((1, 0, 0)))             # only the LAST child moves P
    assert 'SAVEP' not in n


def test_method_chaining_and_registry():
    f = sdf_b200.sphere(1).translate((1, 2, 3)).rotate(0.5).scale(2).shell(0.1)
    assert isinstance(f, sdf_b200.d3.SDF3)
    assert set(['translate', 'rotate', 'orient', 'union', 'shell', 'circular_array', 'twist']) <= set(sdf_b200.d3._ops)
    with pytest.raises(AttributeError):
        sdf_b200.sphere(1).no_such_op
    assert getattr(sdf_b200.sphere(1), '_k', None) is None


def test_reference_quirks_are_kept():
    a = sdf_b200.sphere(1)
    assert sdf_b200.rotate_to(a, (0, 0, 1), (0, 0, 2)) is a               # dot == 1 -> identity (d3.py:366-367)
    assert sdf_b200.slab(z0=0)._node.kind == 'fold'                        # slab returns the inner intersection
    b = sdf_b200.box(a=(0, 0, 0), b=(1, 2, 3))
    assert b._node.consts == [0.5, 1.0, 1.5, 0.5, 1.0, 1.5]


def test_opaque_closure_cannot_be_lowered():
    @sdf_b200.sdf3
    def blob(r):
        def f(p):
            return np.linalg.norm(p, axis=1) - r
        return f
    s = blob(2.0)
    assert float(s(np.array([[3.0, 0, 0]]))[0, 0]) == 1.0                  # still callable in numpy
    with pytest.raises(ir.NotLowerableError):
        ir.lower(s | sdf_b200.sphere(1))


def test_program_size_limit():
    f = sdf_b200.sphere(0.1)
    for i in range(300):
        f = f | sdf_b200.sphere(0.1).translate((i, 0, 0))
    prog = ir.lower(f)                                  # a left-leaning chain shares ONE point slot
    assert prog.pslots == 1 and len(prog.code) > 600
    g = sdf_b200.sphere(0.1).translate((1, 0, 0))
    for i in range(20):                                 # right-leaning: every level needs its own slot
        g = sdf_b200.sphere(0.1).translate((i, 0, 0)) | g.translate((0, 1, 0))
    with pytest.raises(ir.NotLowerableError):
        ir.lower(g)


# ------------------------------------------------------------------ lattice / batching

def test_lattice_matches_reference_construction():
    f = sdf_b200.sphere(1)
    b, step, (X, Y, Z) = core.make_lattice(f, bounds=((-1.25,) * 3, (1.25,) * 3), step=2.5 / 64)
    assert len(X) == len(Y) == len(Z) == 64 and X[0] == -1.25
    Xr, Yr, Zr = sdf_ref.lattice(((-1.25,) * 3, (1.25,) * 3), step=2.5 / 64)
    assert np.array_equal(X, Xr)
    # samples -> isotropic step (core.py:96-98)
    _, (dx, dy, dz), _ = core.make_lattice(f, bounds=((0, 0, 0), (1, 2, 4)), samples=8 ** 3)
    assert dx == dy == dz == (8.0 / 512) ** (1 / 3)


@pytest.mark.parametrize('n,bs', [(64, 32), (65, 32), (33, 32), (100, 32), (96, 40), (5, 8)])
def test_batch_counts_match_reference_slicing(n, bs):
    X = np.arange(n, dtype=np.float64)
    (nbx, nby, nbz), samples = core.batch_counts(X, X, X, bs)
    ref = sdf_ref.batches_of(X, X, X, bs)
    assert nbx * nby * nbz == len(ref)
    assert samples == sum(len(a) * len(b) * len(c) for a, b, c in ref)


def test_partition_layers_properties():
    rng = np.random.default_rng(0)
    for _ in range(200):
        n, world = int(rng.integers(1, 40)), int(rng.integers(1, 9))
        a = rng.integers(0, 50, n) * (rng.random(n) < 0.6)
        parts = multi.partition_layers(a, world)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        assert all(b <= e for b, e in parts)
    parts = multi.partition_layers([10] * 16, 8)
    assert parts == [(2 * i, 2 * i + 2) for i in range(8)]
    sums = [sum([0, 0, 5, 10, 10, 5, 0, 0][b:e]) for b, e in multi.partition_layers([0, 0, 5, 10, 10, 5, 0, 0], 2)]
    assert sums == [15, 15]


def test_tau_scales_with_magnitude():
    p = ir.lower(sdf_b200.sphere(1))
    X = np.linspace(-1, 1, 8)
    t1 = core.mixed_tau(p, X, X, X)
    t2 = core.mixed_tau(p, X * 100, X, X)
    assert 0 < t1 < 1e-4 and abs(t2 / t1 - 100) < 1e-9


# ------------------------------------------------------------------ sinks

def test_binary_stl_layout(tmp_path):
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1]], np.float64)
    path = tmp_path / 'a.stl'
    stl.write_binary_stl(str(path), pts)
    raw = path.read_bytes()
    assert len(raw) == 80 + 4 + 2 * 50 and raw[:80] == b'\x00' * 80
    assert struct.unpack('<I', raw[80:84])[0] == 2
    rec = struct.unpack('<12fH', raw[84:134])
    assert rec[:3] == (0.0, 0.0, 1.0) and rec[3:12] == (0, 0, 0, 1, 0, 0, 0, 1, 0) and rec[12] == 0


def test_generate_verbose_lines(monkeypatch, capsys):
    """The verbose output keeps the reference's format strings (core.py:105-108,125-126,145-148)."""
    fake_info = dict(triangles=2, skipped=3, empty=4, nonempty=1)

    def fake_run(sdf, X, Y, Z, batch_size, sparse, precision, **kw):
        return np.zeros((2, 9), np.float32), dict(fake_info)
    monkeypatch.setattr(core, 'run_batches', fake_run)
    pts = core.generate(sdf_b200.sphere(1), step=0.5, bounds=((-1, -1, -1), (1, 1, 1)), workers=7)
    out = capsys.readouterr().out
    assert 'min -1, -1, -1' in out and 'max 1, 1, 1' in out and 'step 0.5, 0.5, 0.5' in out
    assert '64 samples in 1 batches with 7 workers' in out
    assert '3 skipped, 4 empty, 1 nonempty' in out and re.search(r'2 triangles in [0-9.e+-]+ seconds', out)
    assert pts.shape == (6, 3)


def test_drop_in_package_names():
    import sdf
    for name in ['sphere', 'box', 'cylinder', 'torus', 'union', 'intersection', 'difference', 'translate', 'rotate',
                 'scale', 'orient', 'shell', 'generate', 'save', 'sample_slice', 'show_slice', 'write_binary_stl',
                 'pi', 'degrees', 'radians', 'X', 'Y', 'Z', 'ORIGIN', 'UP', 'sdf3', 'op3', 'op32', 'sdf2', 'op2',
                 'op23', 'SDF3', 'SDF2', 'np']:
        assert hasattr(sdf, name), name
    assert sdf.X.shape == (3,)                    # 3-D constants win over the 2-D ones (__init__.py:5-7)


# ------------------------------------------------------------------ C-ABI surface

def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'sdfb200.h')).read()
    declared = set(re.findall(r'\b(sdfb200_[a-z0-9_]+)\s*\(', header)) - {'sdfb200_prog'}
    assert declared == set(_abi.EXPORTS), declared ^ set(_abi.EXPORTS)
    lib = ctypes.CDLL(_abi.LIB_PATH)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.sdfb200_abi_version() == 2          # v2: timings_ms[8], sdfb200_generate_window


def test_no_gpu_means_loud_failure_not_cpu_fallback():
    """On a box without a CUDA device the product path must raise (never route through the oracle)."""
    if _abi.device_count() > 0:
        pytest.skip('a GPU is present')
    with pytest.raises(_abi.SdfB200Error) as ei:
        sdf_b200.sphere(1).generate(step=0.5, bounds=((-1, -1, -1), (1, 1, 1)), verbose=False)
    assert ei.value.code == _abi.E_NOGPU and 'no CPU fallback' in str(ei.value)
    with pytest.raises(_abi.SdfB200Error):
        sdf_b200.sphere(1)(np.zeros((1, 3)))


def test_runtime_specialisation_compiles_without_a_gpu():
    """The generated straight-line evaluator of a tree compiles with NVRTC for sm_100a (no GPU needed)."""
    from sdf_b200 import core
    ms = models.widened_corpus(sdf_b200)
    for name, kernel in (('weave', 4), ('pawn', 6), ('slice', 7), ('x_polygon', 0), ('extrude_to', 8)):
        h, _ = core.program_handle(ms[name])
        assert h.jit_probe(kernel) > 10000, name


def test_star_export_surface_matches_the_reference():
    """sdf/__init__.py:1-27: every public name a `from sdf import *` script may rely on exists; text / image are
    host nodes that fail like the reference's on a missing font / file (Pillow's OSError), Mesh.sdf (pyopenvdb) raises."""
    import sdf
    for name in ('np', 'pi', 'degrees', 'radians', 'X', 'Y', 'Z', 'ORIGIN', 'UP', 'ease', 'd2', 'd3', 'Mesh', 'text', 'image',
                 'measure_text', 'measure_image', 'generate', 'save', 'sample_slice', 'show_slice', 'write_binary_stl',
                 'sdf2', 'sdf3', 'op2', 'op3', 'op23', 'op32', 'SDF2', 'SDF3', 'sphere', 'box', 'circle', 'union', 'repeat'):
        assert hasattr(sdf, name), name
    assert sdf.X is sdf_b200.d3.X and len(sdf.X) == 3            # the 3-D constants win (sdf/__init__.py:5-7)
    for call in (lambda: sdf.text('no-such-font', 'x'), lambda: sdf.image('no-such-file.png'),
                 lambda: sdf.measure_text('no-such-font', 'x')):
        with pytest.raises(OSError):
            call()
    with pytest.raises(NotImplementedError):
        sdf.Mesh(np.zeros((3, 3)), np.array([[0, 1, 2]])).sdf(0.25)


def test_sample_grouping_plan_follows_the_tree_structure():
    """sdfb200_jit_plan: the tile kernel runs a thread's samples along the axis most of the tree does not depend on.
    The `exact` kernel (sign masks by ballot) always uses thread-runs; the value-tile kernel of the other precisions
    falls back to its flat 4-sample mapping when there is too little to share."""
    from sdf_b200 import core
    S = sdf_b200
    cases = {
        'gearlike': (models.CONFIGS['gearlike'](S), 2, 11),                  # cylinders + circular array about Z
        # rotate_to's cos(pi/2) is 6e-17, not 0: the analysis (and the float32 code) must see through it
        'array_x': (S.cylinder(0.25).circular_array(8, 1).orient(S.X) & S.sphere(2), 0, 11),
        'array_y': (S.cylinder(0.25).circular_array(8, 1).orient(S.Y) & S.sphere(2), 1, 11),
        'extrusion': (S.polygon([(0, 0), (1, 0), (1.2, 0.8), (0.5, 1.3), (-0.2, 0.7)]).extrude(2) - S.circle(0.4).extrude(3), 2, 11),
        # too little to save: any axis, any run length (exact); the flat 4-sample mapping (mixed, measured faster there)
        'sphere': (S.sphere(1), -1, 4),
        'small': (S.cylinder(1).orient(S.X) & S.sphere(2), -1, 4),
        'csg': (models.CONFIGS['csg'](S), -1, 4),
    }
    for name, (f, axis, group) in cases.items():
        h, _ = core.program_handle(f)
        got_axis, got_group, frac = h.jit_plan(3)
        assert 0.0 <= frac <= 1.0
        if axis >= 0:
            assert (got_axis, got_group) == (axis, group), (name, got_axis, got_group)
            assert frac > 0.3, (name, frac)
        else:
            assert got_axis in (0, 1, 2) and got_group in (11, 6, 3), (name, got_axis, got_group)
            assert h.jit_plan(1)[:2] == (-1, 4), (name, h.jit_plan(1))
    # a deep tree needs more registers per sample: the group shrinks until ptxas stops spilling
    h, _ = core.program_handle(models.widened_corpus(S)['weave'])
    assert h.jit_plan(3)[1] < 11
    # forced grouping (A/B timing, tests)
    import os
    os.environ['SDFB200_JIT_SHARE'] = '1'; os.environ['SDFB200_JIT_N'] = '4'
    try:
        assert _abi.ProgramHandle(ir.lower(cases['gearlike'][0])).jit_plan(3)[:2] == (1, 4)
    finally:
        os.environ.pop('SDFB200_JIT_SHARE'); os.environ.pop('SDFB200_JIT_N')


def test_jit_defines_hook_compiles():
    """SDFB200_JIT_DEFINES (extra #defines of the run-time compiled unit, an A/B hook for experiments) is part of
    the source: the kernels keep compiling with it; unset, the generated source is untouched."""
    import os
    prog = ir.lower(models.CONFIGS['gearlike'](sdf_b200))
    os.environ['SDFB200_JIT_DEFINES'] = 'SDF_AB_MARKER SDF_AB_VALUE=1'
    try:
        h = _abi.ProgramHandle(prog)
        sizes = [h.jit_probe(k) for k in (0, 4, 7, 8)]
    finally:
        os.environ.pop('SDFB200_JIT_DEFINES')
    assert all(sz > 10000 for sz in sizes)

def test_specialised_kernels_are_cached_on_disk(tmp_path):
    import os, time
    old = os.environ.get('SDFB200_JIT_CACHE')
    os.environ['SDFB200_JIT_CACHE'] = str(tmp_path / 'jit')
    try:
        prog = ir.lower(sdf_b200.torus(1, 0.25).twist(0.5) | sdf_b200.box(0.7))
        t0 = time.time(); a = _abi.ProgramHandle(prog).jit_probe(4); t1 = time.time()
        files = sorted(os.listdir(tmp_path / 'jit'))
        assert len(files) == 1 and files[0].endswith('.cubin')
        b = _abi.ProgramHandle(prog).jit_probe(4); t2 = time.time()
        assert a == b and sorted(os.listdir(tmp_path / 'jit')) == files
        assert (t2 - t1) < 0.5 * (t1 - t0)                   # a file read instead of an NVRTC compile
        # a damaged entry is ignored and replaced
        with open(tmp_path / 'jit' / files[0], 'wb') as fh:
            fh.write(b'garbage')
        assert _abi.ProgramHandle(prog).jit_probe(4) == a
        os.environ['SDFB200_JIT_CACHE'] = ''                 # empty: no cache
        assert _abi.ProgramHandle(prog).jit_probe(4) == a
    finally:
        if old is None:
            os.environ.pop('SDFB200_JIT_CACHE', None)
        else:
            os.environ['SDFB200_JIT_CACHE'] = old


def test_large_program_is_accepted_and_small_limits_still_guard():
    f = sdf_b200.sphere(0.1)
    for i in range(12):
        f = f | sdf_b200.sphere(0.1).translate((i, 0, 0)).repeat((50.0, 3.0, 0), padding=1)
    prog = ir.lower(f)
    assert len(prog.code) > ir.IMG_MAX_CODE            # too big for the interpreter image ...
    h = _abi.ProgramHandle(prog)                        # ... but a valid program (runs on the specialised kernels)
    assert h.jit_probe(6) > 10000


def test_program_create_validates():
    lib = _abi.lib()
    h = ctypes.c_void_p()
    bad = np.array([250], np.uint32)                    # unknown opcode
    assert lib.sdfb200_program_create(bad.ctypes.data, 1, None, 0, ctypes.byref(h)) == _abi.E_INVALID
    assert b'unknown opcode' in lib.sdfb200_last_error()
    code = np.array([ir.OPCODE['SPHERE'] | (0 << 16)], np.uint32)
    assert lib.sdfb200_program_create(code.ctypes.data, 1, np.zeros(2).ctypes.data, 2, ctypes.byref(h)) == _abi.E_INVALID
    assert lib.sdfb200_program_create(code.ctypes.data, 1, np.zeros(4).ctypes.data, 4, ctypes.byref(h)) == 0
    lib.sdfb200_program_free(h)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'sdf_b200')):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle|oracle/|oracle\.|libmc_oracle|sdf_ref', src, re.M), fn


def test_product_knows_nothing_of_the_host_execution():
    """tests/hostsim/ (the device code and the library executed on the host) is test infrastructure: no file of the
    product, the bench or the driver entry points mentions it -- there is no route from the product to a CPU path."""
    files = [os.path.join(ROOT, 'bench.py'), os.path.join(ROOT, '__graft_entry__.py'), os.path.join(ROOT, 'include', 'sdfb200.h')]
    for base in ('sdf_b200', 'sdf'):
        for dirpath, _, names in os.walk(os.path.join(ROOT, base)):
            files += [os.path.join(dirpath, fn) for fn in names if fn.endswith(('.py', '.cu', '.cuh', '.h', '.inc'))]
    for path in files:
        assert not re.search(r'hostsim|HOSTSIM', open(path).read()), path


def test_integration_md_stub_binds_the_library():
    """The minimal ctypes stub printed in INTEGRATION.md is executed as written: it loads the library,
    creates a program and reaches sdfb200_generate (which, without a GPU, reports "no CPU fallback")."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, 'INTEGRATION.md')).read()
    block = re.search(r"```python\n# sdf/_b200\.py\n(.*?)```", text, re.S).group(1)
    block = block.replace("ctypes.CDLL('libsdfb200.so')", "ctypes.CDLL(%r)" % os.path.join(root, 'sdf_b200', 'libsdfb200.so'))
    ns = {}
    exec(compile(block, 'INTEGRATION.md', 'exec'), ns)
    prog = ir.lower(sdf_b200.sphere(1))
    h = ns['program'](np.ascontiguousarray(prog.code, np.uint32), np.ascontiguousarray(prog.consts, np.float64))
    assert h.value
    X = np.arange(-1.5, 1.5, 0.1)
    if _abi.device_count() > 0:
        pts, stats = ns['generate'](h, X, X, X, capacity=1 << 16)
        assert len(pts) > 0 and len(pts) % 3 == 0
    else:
        with pytest.raises(RuntimeError) as ei:
            ns['generate'](h, X, X, X, capacity=1 << 10)
        assert 'no CPU fallback' in str(ei.value)


def test_sample_slice_host_logic_matches_reference(golden_dir):
    """core.sample_slice (core.py:202-232): point order, reshape, extent and axis names -- the evaluator is
    replaced by the oracle here (any callable works once bounds are given), the vectors come from the
    REAL reference."""
    from oracle import sdf_ref
    from sdf_b200 import core
    g = np.load(golden_dir + '/engine.npz')
    f = models.CONFIGS['csg'](sdf_b200)
    cpu = lambda P: sdf_ref.evaluate(f, P)
    sb = ((-1.0, -0.9, -0.8), (1.1, 1.0, 0.7))
    for key, kw in (('x', dict(x=0.13)), ('y', dict(y=-0.2)), ('z', dict(z=0.31))):
        a, extent, axes = core.sample_slice(cpu, w=24, h=17, bounds=sb, **kw)
        assert a.shape == (24, 17) and np.array_equal(a, g['slice_' + key])
        assert np.array_equal(np.array(extent, np.float64), g['slice_extent_' + key])
        assert axes == ''.join(chr(c) for c in g['slice_axes_' + key])
    with pytest.raises(Exception):
        core.sample_slice(cpu, bounds=sb)


def test_binary_stl_bytes_equal_the_reference_writer(golden_dir, tmp_path):
    """stl.write_binary_stl against the file the REAL reference wrote for the same soup (float32 normals
    computed operation by operation like numpy, a zero-area triangle's NaN normal included)."""
    g = np.load(golden_dir + '/engine.npz')
    path = tmp_path / 'a.stl'
    with np.errstate(all='ignore'):
        stl.write_binary_stl(str(path), g['stl_points'])
    assert path.read_bytes() == g['stl_bytes'].tobytes()
    with np.errstate(all='ignore'):
        stl.write_binary_stl(str(path), g['stl_points'].astype(np.float32).reshape(-1, 3))     # what generate() returns
    assert path.read_bytes() == g['stl_bytes'].tobytes()


def test_guard_band_follows_the_amplification_of_the_tree():
    """errbound.gain: the a-priori factor core.mixed_tau multiplies its guard band with.  The BASELINE
    configurations do not amplify float32 rounding (gain ~ 1: their tau -- and their speed -- is what was
    validated); a warp that divides by r1 - r0 = 1e-3, a huge twist rate, a steep easing widen the band; a
    tree with a jump (square easing) sends every sample to float64."""
    import math
    from sdf_b200 import core, errbound
    S = sdf_b200
    for name, build in models.CONFIGS.items():
        g = errbound.gain(ir.lower(build(S)), 2.75)
        assert 1.0 <= g < 1.1, (name, g)
    assert errbound.gain(ir.lower(S.box(1).bend_radial(1.0, 1.001, 0.5)), 2.5) > 50
    assert errbound.gain(ir.lower(S.box(1).twist(200.0)), 2.5) > 20
    assert errbound.gain(ir.lower(S.sphere(1).scale(0.001)), 2.5) < 2          # undone by its own MULC
    mild = errbound.gain(ir.lower(S.box(1).bend_linear(S.X * -1, S.X, S.Z * 0.5, S.ease.in_quad)), 2.5)
    steep = errbound.gain(ir.lower(S.box(1).bend_linear(S.X * -1, S.X, S.Z * 0.5, S.ease.in_out_elastic)), 2.5)
    assert 1.0 < mild < steep, (mild, steep)
    assert math.isinf(errbound.gain(ir.lower(S.box(1).bend_linear(S.X * -1, S.X, S.Z * 0.5, S.ease.in_square)), 2.5))
    # every opcode has a rule (an unknown one would return inf and silently push everything to float64)
    for name, f in models.build(S).items():
        if 'square' in name or 'circ' in name:
            continue
        assert math.isfinite(errbound.gain(ir.lower(f), 2.5)), name
    # mixed_tau = heuristic x gain
    X = np.arange(-2.5, 2.5, 0.1)
    f = S.box(1).bend_radial(1.0, 1.001, 0.5)
    h, _ = core.program_handle(f)
    h0, _ = core.program_handle(S.box(1).bend_radial(1.0, 2.0, 0.5))
    assert core.mixed_tau(h.program, X, X, X) > 30 * core.mixed_tau(h0.program, X, X, X)


def test_user_closures_compose_with_the_library_combinators():
    """README "writing your own SDFs": a tree that holds Python closures evaluates piecewise -- closures on the host, as
    the reference calls them, combined by the reference's CSG / transform formulas (core._compose_around_closure).
    Checked against the oracle on the equivalent library tree (all-closure trees need no GPU)."""
    from oracle import sdf_ref
    S = sdf_b200

    @S.sdf3
    def my_sphere(r):
        def f(p):
            return np.linalg.norm(p, axis=1) - r
        return f

    @S.sdf3
    def my_box(h):
        def f(p):
            q = np.abs(p) - h
            return np.linalg.norm(np.maximum(q, 0), axis=1) + np.minimum(np.max(q, axis=1), 0)
        return f

    a, b = my_sphere(0.8), my_box(0.6)
    la, lb = S.sphere(0.8), S.box(1.2)
    P = models.random_points(2000, seed=5)
    cases = [
        (a | b, la | lb), (a & b, la & lb), (a - b, la - lb),
        (a.union(b, k=0.2), la.union(lb, k=0.2)), (a.difference(b.k(0.15)), la.difference(lb.k(0.15))),
        (a.intersection(b, k=0.1), la.intersection(lb, k=0.1)), (a.blend(b, k=0.3), la.blend(lb, k=0.3)),
        (a.negate(), la.negate()), (b.dilate(0.1), lb.dilate(0.1)), (b.erode(0.05), lb.erode(0.05)), (a.shell(0.1), la.shell(0.1)),
        (b.translate((0.2, -0.1, 0.3)), lb.translate((0.2, -0.1, 0.3))), (b.rotate(0.7, (1, 2, 3)), lb.rotate(0.7, (1, 2, 3))),
        (b.scale((1.5, 0.5, 2.0)), lb.scale((1.5, 0.5, 2.0))), ((a | b.translate(S.X * 0.5)).orient(S.Y) - a.scale(0.5), (la | lb.translate(S.X * 0.5)).orient(S.Y) - la.scale(0.5)),
    ]
    for k, (mine, lib) in enumerate(cases):
        got = mine(P).reshape(-1)
        ref = sdf_ref.evaluate(lib, P).reshape(-1)
        assert np.max(np.abs(got - ref)) <= 1e-12, (k, np.max(np.abs(got - ref)))
    with pytest.raises(ir.NotLowerableError):
        a.twist(1.0)(P)                       # a warp above a closure: not one of the composition nodes
    with pytest.raises(ir.NotLowerableError):
        ir.lower(a | la)


def test_plain_c_client_binds_the_abi_and_fails_loudly_without_a_gpu(tmp_path):
    """include/sdfb200.h is C99 (no C++, no CUDA, no torch types), a C program links the library, malformed input is
    refused with a message, and without a device the generate call returns SDFB200_E_NOGPU -- no CPU fallback."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present: tests/test_gpu_zzz_cabi_client.py runs the client for real')
    from util_cabi import build_c_client
    exe = build_c_client(tmp_path)
    out = subprocess.run([exe, '64', '32'], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith('NOGPU error=') and 'no CPU fallback' in out.stdout, out.stdout

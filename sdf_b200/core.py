"""Mesh-generation engine (host side) -- the drop-in for the reference's sdf/core.py.

``generate`` keeps the reference signature (core.py:84-88) and its verbose
output (core.py:105-108, 124-126, 145-148).  What used to be a ThreadPool of
numpy/skimage workers (core.py:131-141) is one call through the C-ABI
(``sdfb200_generate``): skip pre-pass, fused SDF evaluation, marching cubes and
world mapping all run on the GPU.  ``workers`` is accepted and ignored.

Lattice construction stays on the host in float64 with ``np.arange`` so the
sample coordinates are bit-identical to the reference's (core.py:110-112).
"""
import multiprocessing
import os
import time

import numpy as np

from . import _abi, progress, stl
from .ir import NotLowerableError, lower

WORKERS = multiprocessing.cpu_count()
SAMPLES = 2 ** 22
BATCH_SIZE = 32

TAU_ULPS = 64.0          # guard band of the float32 pass, in float32 ulps of the largest magnitude in play
TAU_ULPS_PER_WORD = 4.0  # ... plus this many per program word (rounding accumulates along the tree)

_EPS32 = float(np.finfo(np.float32).eps)
_handles = {}            # (code bytes, consts bytes) -> ProgramHandle
_capacity_hint = {}      # (program key, lattice dims, batch, sparse) -> triangles needed last time


def default_device():
    for var in ('SDFB200_DEVICE', 'LOCAL_RANK'):
        v = os.environ.get(var)
        if v is not None:
            return int(v)
    return 0


def program_handle(sdf):
    """The lowered program of ``sdf`` and its device handle.  A tree is immutable except for ``.k()`` set after
    construction (honoured like the reference does, dn.py:12): the lowering is cached on the root object and
    redone only when some ``_k`` has changed since (ir.MUTATIONS)."""
    from . import ir
    cached = sdf.__dict__.get('_lowered')
    if cached is not None and cached[0] == ir.MUTATIONS[0]:
        return cached[1], cached[2]
    h, key = _program_handle(sdf)
    try:
        object.__setattr__(sdf, '_lowered', (ir.MUTATIONS[0], h, key))
    except AttributeError:
        pass
    return h, key


def _program_handle(sdf):
    prog = lower(sdf)
    key = (prog.code.tobytes(), prog.consts.tobytes())
    h = _handles.get(key)
    if h is None:
        if len(_handles) > 64:
            _handles.clear()
        h = _abi.ProgramHandle(prog)
        _handles[key] = h
    return h, key


def _lowerable(obj):
    try:
        program_handle(obj)
        return True
    except NotLowerableError:
        return False


def _compose_around_closure(obj, p):
    """``obj(p)`` for a tree that holds a user-written numpy closure somewhere below ``obj``.

    A Python closure can only run on the host, so the nodes ABOVE it cannot be fused into the CUDA evaluator
    either.  Every maximal sub-tree without a closure still evaluates on the GPU (one call each); the closure is
    called as the reference calls it (d3.py:24-25); and the handful of composition nodes people put around their
    own SDFs and around text() / image() -- the CSG folds, blend, negate / dilate / erode / shell, translate /
    scale / rotate in 2-D and 3-D, extrude / extrude_to / revolve (dn.py:7-80, d3.py:329-360, d2.py:211-287) --
    combine those fields on the host with the reference's formulas.  Any other node above a
    closure raises NotLowerableError.  Lowered trees (everything the library itself builds) never come here."""
    from .ir import node_of
    from . import d2
    n = node_of(obj)
    if isinstance(obj, d2.SDF2) and p.shape[1] == 3:
        p = p[:, :2]
    if n.kind == 'opaque':
        return np.asarray(obj.f(p), np.float64).reshape(-1)
    if _lowerable(obj):
        return np.asarray(obj(p), np.float64).reshape(-1)         # a closure-free sub-tree: one GPU evaluation
    kids = n.children
    if n.kind == 'fold' or n.kind == 'blend':
        d1 = _compose_around_closure(kids[0], p)
        for b in kids[1:]:
            d2 = _compose_around_closure(b, p)
            K = n.extra or getattr(b, '_k', None)
            if n.kind == 'blend':                                        # dn.py:52-60
                d1 = K * d2 + (1 - K) * d1
            elif n.op == 'union':                                        # dn.py:7-20
                if K is None:
                    d1 = np.minimum(d1, d2)
                else:
                    h = np.clip(0.5 + 0.5 * (d2 - d1) / K, 0, 1)
                    d1 = d2 + (d1 - d2) * h - K * h * (1 - h)
            elif n.op == 'difference':                                   # dn.py:22-35
                if K is None:
                    d1 = np.maximum(d1, -d2)
                else:
                    h = np.clip(0.5 - 0.5 * (d2 + d1) / K, 0, 1)
                    d1 = d1 + (-d2 - d1) * h + K * h * (1 - h)
            else:                                                        # dn.py:37-50
                if K is None:
                    d1 = np.maximum(d1, d2)
                else:
                    h = np.clip(0.5 - 0.5 * (d2 - d1) / K, 0, 1)
                    d1 = d2 + (d1 - d2) * h + K * h * (1 - h)
        return d1
    if n.kind == 'value' and n.op in ('NEGATE', 'ADDC', 'SHELL'):        # dn.py:62-80
        d = _compose_around_closure(kids[0], p)
        return -d if n.op == 'NEGATE' else (d + n.consts[0] if n.op == 'ADDC' else np.abs(d) - n.consts[0] / 2)
    dim = p.shape[1]
    if n.kind == 'transform' and n.op == 'TRANSLATE':                    # d3.py:329-333, d2.py:211-215
        return _compose_around_closure(kids[0], p - np.array(n.consts[:dim]))
    if n.kind == 'transform' and n.op == 'ROTATE':                       # d3.py:347-360, d2.py:228-238
        return _compose_around_closure(kids[0], np.dot(p, np.array(n.consts).reshape(3, 3)[:dim, :dim]))
    if n.kind == 'scale':                                                # d3.py:335-345, d2.py:217-226
        return _compose_around_closure(kids[0], p / np.array(n.consts[:dim])) * n.consts[3]
    if n.kind == 'extrude' or n.op == 'EXTRUDE_TO':                      # d2.py:261-279
        if n.kind == 'extrude':
            h = n.consts[0]
            d = _compose_around_closure(kids[0], p[:, :2])
        else:
            (_, (h,)), ease_id = n.extra[0], n.extra[1]
            d1 = _compose_around_closure(kids[0], p[:, :2])
            d2_ = _compose_around_closure(kids[1], p[:, :2])
            d = d1 + (d2_ - d1) * _ease_by_id(ease_id)(np.clip(p[:, 2] / h, -0.5, 0.5) + 0.5)
        w = np.stack([d, np.abs(p[:, 2]) - h / 2], axis=-1)
        return np.minimum(np.maximum(w[:, 0], w[:, 1]), 0) + np.linalg.norm(np.maximum(w, 0), axis=1)
    if n.kind == 'transform' and n.op == 'REVOLVE':                      # d2.py:281-287
        q = np.stack([np.linalg.norm(p[:, :2], axis=1) - n.consts[0], p[:, 2]], axis=-1)
        return _compose_around_closure(kids[0], q)
    raise NotLowerableError('node %r sits above a Python closure: only CSG folds, blend, negate / dilate / erode / shell, '
                            'translate / scale / rotate and extrude / extrude_to / revolve can be combined with '
                            'user-written SDFs and text() / image()' % (n.label,))


def _ease_by_id(ease_id):
    from . import ease
    for name, i in ease.EASE_ID.items():
        if i == ease_id:
            return getattr(ease, name)
    raise ValueError('unknown easing id %r' % (ease_id,))


def evaluate(sdf, p, device=None):
    """``sdf(p)``: float64 points in, float64 distances out (d3.py:24-25), on the GPU.

    numpy in -> numpy out (staged through device memory inside the C-ABI call);
    a CUDA torch tensor in -> a CUDA tensor of the same dtype out (float32 uses
    the streamed float32 evaluator).  A tree that holds a user-written numpy closure is evaluated piecewise
    (``_compose_around_closure``).
    """
    try:
        h, _ = program_handle(sdf)
    except NotLowerableError:
        if hasattr(p, 'data_ptr'):
            raise
        pts = np.ascontiguousarray(p, dtype=np.float64)
        return _compose_around_closure(sdf, pts.reshape(-1, pts.shape[-1]))
    if isinstance(p, np.ndarray) or not hasattr(p, 'data_ptr'):
        pts = np.ascontiguousarray(p, dtype=np.float64).reshape(-1, 3)
        out = np.empty(len(pts), np.float64)
        h.eval_f64(pts, out, len(pts), device=default_device() if device is None else device)
        return out
    import torch
    pts = p.contiguous().reshape(-1, 3)
    dev = pts.device.index if pts.is_cuda else (default_device() if device is None else device)
    stream = torch.cuda.current_stream(dev).cuda_stream if pts.is_cuda else None
    out = torch.empty(len(pts), dtype=pts.dtype, device=pts.device)
    if pts.dtype == torch.float32:
        h.eval_f32(pts, out, len(pts), device=dev, stream=stream)
    elif pts.dtype == torch.float64:
        h.eval_f64(pts, out, len(pts), device=dev, stream=stream)
    else:
        raise TypeError('points must be float32 or float64')
    return out


def _cartesian_product(*arrays):
    grids = np.meshgrid(*arrays, indexing='ij')
    return np.stack([g.reshape(-1) for g in grids], axis=-1)


def _estimate_bounds(sdf):
    """core.py:62-82: shrink a 16^3 probe lattice from +-1e9 around |d| <= half a cell diagonal."""
    s = 16
    lo = np.full(3, -1e9)
    hi = np.full(3, 1e9)
    prev = None
    for _ in range(32):
        axes = [np.linspace(lo[k], hi[k], s) for k in range(3)]
        d = np.array([a[1] - a[0] for a in axes])
        threshold = np.linalg.norm(d) / 2
        if threshold == prev:
            break
        prev = threshold
        P = _cartesian_product(*axes)
        volume = sdf(P).reshape((s, s, s))
        where = np.argwhere(np.abs(volume) <= threshold)
        x1, y1, z1 = (lo[0], lo[1], lo[2]) + where.max(axis=0) * d + d / 2
        x0, y0, z0 = (lo[0], lo[1], lo[2]) + where.min(axis=0) * d - d / 2
        lo = np.array([x0, y0, z0])
        hi = np.array([x1, y1, z1])
    return ((lo[0], lo[1], lo[2]), (hi[0], hi[1], hi[2]))


def make_lattice(sdf, step=None, bounds=None, samples=SAMPLES):
    """Bounds, step and the three float64 axes exactly as core.py:94-112 builds them."""
    if bounds is None:
        bounds = _estimate_bounds(sdf)
    (x0, y0, z0), (x1, y1, z1) = bounds
    if step is None and samples is not None:
        volume = (x1 - x0) * (y1 - y0) * (z1 - z0)
        step = (volume / samples) ** (1 / 3)
    try:
        dx, dy, dz = step
    except TypeError:
        dx = dy = dz = step
    X = np.arange(x0, x1, dx)
    Y = np.arange(y0, y1, dy)
    Z = np.arange(z0, z1, dz)
    return ((x0, y0, z0), (x1, y1, z1)), (dx, dy, dz), (X, Y, Z)


def batch_counts(X, Y, Z, batch_size):
    """Number of batches per axis and total sample count incl. shared planes (core.py:114-122)."""
    def lens(n):
        return [min(batch_size + 1, n - i) for i in range(0, n, batch_size)]
    lx, ly, lz = lens(len(X)), lens(len(Y)), lens(len(Z))
    num_samples = sum(lx) * sum(ly) * sum(lz)
    return (len(lx), len(ly), len(lz)), num_samples


def mixed_tau(program, X, Y, Z):
    """Guard band of the float32 pass (precisions `mixed` and `exact`): samples with |d| < tau are re-evaluated
    in float64.  TAU_ULPS float32 ulps of the largest coordinate / constant (+ a few per program word) -- tuned
    and tested on trees whose constants do not amplify rounding errors -- times the a-priori amplification of
    THIS tree (errbound.gain: reciprocals of small constant differences, steep easings, large twist rates ...;
    1 for the BASELINE configurations, inf -- float64 everywhere -- for a tree with a jump)."""
    from . import errbound
    cmax = 1.0
    for a in (X, Y, Z):
        if len(a):
            cmax = max(cmax, abs(float(a[0])), abs(float(a[-1])))
    mag = max(cmax, min(program.scale_hint, 1e6))
    ulps = TAU_ULPS + TAU_ULPS_PER_WORD * len(program.code)
    cache = program.__dict__.setdefault('_gain', {})          # per program and coordinate magnitude
    key = float(np.float32(cmax))
    g = cache.get(key)
    if g is None:
        g = cache[key] = errbound.gain(program, cmax)
    return ulps * _EPS32 * mag * g


def default_capacity(X, Y, Z):
    nx, ny, nz = len(X), len(Y), len(Z)
    return int(max(1 << 16, 8 * (nx * ny + ny * nz + nx * nz)))


def run_batches(sdf, X, Y, Z, batch_size=BATCH_SIZE, sparse=True, precision='exact',
                zb=(0, -1), device=None, out=None, stream=None):
    """One C-ABI generate call (with the retry on capacity overflow).

    ``out``: None -> a fresh host ndarray is returned; or a caller-owned torch CUDA tensor /
    numpy array of shape (capacity, 9) float32.  Returns (triangles (T,9) view, info dict).
    """
    h, key = program_handle(sdf)
    X = np.ascontiguousarray(X, np.float64)
    Y = np.ascontiguousarray(Y, np.float64)
    Z = np.ascontiguousarray(Z, np.float64)
    tau = mixed_tau(h.program, X, Y, Z)
    dev = default_device() if device is None else device
    hint_key = (key, len(X), len(Y), len(Z), batch_size, bool(sparse), tuple(zb))
    own = out is None
    cap = _capacity_hint.get(hint_key, default_capacity(X, Y, Z)) if own else int(out.shape[0])
    while True:
        if own:
            # page-locked and device-mapped: the kernel writes the triangles straight into it
            out = _abi.pinned_pool.take(cap * 9)[:cap * 9].reshape(cap, 9)
        try:
            n_tri, stats, timings = h.generate(X, Y, Z, batch_size, sparse, precision, tau, out, cap,
                                               zb=zb, device=dev, stream=stream)
            break
        except _abi.CapacityError as e:
            if not own:
                raise
            cap = int(e.needed * 1.05) + 1024
    _capacity_hint[hint_key] = int(n_tri * 1.05) + 1024
    info = dict(triangles=n_tri, skipped=int(stats[0]), empty=int(stats[1]), nonempty=int(stats[2]),
                evaluated=int(stats[3]), refined=int(stats[4]), active_cubes=int(stats[5]),
                prepass_ms=float(timings[0]), fused_ms=float(timings[1]), device_ms=float(timings[2]),
                d2h_ms=float(timings[3]), tile_eval_ms=float(timings[4]), refine_ms=float(timings[5]),
                extract_ms=float(timings[6]), classify_ms=float(timings[7]), tau=tau, precision=precision)
    return out[:n_tri], info


def _skip_host(sdf, job):
    """core.py:28-43 on the host (only for trees the CUDA evaluator cannot lower): |d(centre)| against the half
    diagonal, then the eight corners in itertools.product order."""
    import itertools
    X, Y, Z = job
    x0, x1, y0, y1, z0, z1 = X[0], X[-1], Y[0], Y[-1], Z[0], Z[-1]
    x, y, z = (x0 + x1) / 2, (y0 + y1) / 2, (z0 + z1) / 2
    r = abs(sdf(np.array([(x, y, z)])).reshape(-1)[0])
    d = np.linalg.norm(np.array((x - x0, y - y0, z - z0)))
    if r <= d:
        return False
    corners = np.array(list(itertools.product((x0, x1), (y0, y1), (z0, z1))))
    values = sdf(corners).reshape(-1)
    same = np.all(values > 0) if values[0] > 0 else np.all(values < 0)
    return bool(same)


def _generate_opaque(sdf, X, Y, Z, batch_size, sparse, device):
    """The route for a tree that holds a user-written numpy closure (README "writing your own SDFs"; ir.py kind
    'opaque'): the closure cannot run on the GPU, so each batch is evaluated by calling the tree exactly like the
    reference does (core.py:45-52; library nodes below the closure still evaluate on the GPU through their own
    ``__call__``), and the batch volume goes through marching cubes + world mapping on the device
    (``sdfb200_mc_volume``).  Same batches, same _skip rule, same float32 cast as core.py / skimage; a batch of
    more than 33 samples per axis is meshed in pieces (each mapped to world space from its own first sample, which
    can differ from the reference's batch-relative mapping in the last float32 bit)."""
    import itertools
    h, _ = program_handle(_context_sdf())
    dev = default_device() if device is None else device
    s = batch_size
    Xs = [X[i:i + s + 1] for i in range(0, len(X), s)]
    Ys = [Y[i:i + s + 1] for i in range(0, len(Y), s)]
    Zs = [Z[i:i + s + 1] for i in range(0, len(Z), s)]
    skipped = empty = nonempty = evaluated = 0
    out = []
    for job in itertools.product(Xs, Ys, Zs):
        if sparse and _skip_host(sdf, job):
            skipped += 1
            continue
        P = _cartesian_product(*job)
        volume = np.asarray(sdf(P), np.float64).reshape((len(job[0]), len(job[1]), len(job[2])))
        evaluated += volume.size
        found = 0
        # the device meshes volumes of up to 33^3 samples: a larger batch goes in pieces that share a plane --
        # marching cubes is per cube, so the pieces' soups add up to the batch's
        for ix, iy, iz in itertools.product(*(range(0, max(len(a) - 1, 1), 32) for a in job)):
            piece = volume[ix:ix + 33, iy:iy + 33, iz:iz + 33]
            tris = h.mc_volume(piece, job[0][ix:ix + 33], job[1][iy:iy + 33], job[2][iz:iz + 33], device=dev)
            if len(tris):
                found += 1
                out.append(tris)
        if found:
            nonempty += 1
        else:
            empty += 1
    tris = np.concatenate(out) if out else np.zeros((0, 9), np.float32)
    info = dict(triangles=len(tris), skipped=skipped, empty=empty, nonempty=nonempty, evaluated=evaluated, refined=0,
                active_cubes=0, precision='f64 on the host (Python closure in the tree) + device marching cubes')
    return tris, info


_ctx_sdf = []


def _context_sdf():
    if not _ctx_sdf:
        from . import d3
        _ctx_sdf.append(d3.sphere(1))
    return _ctx_sdf[0]


def generate(
        sdf,
        step=None, bounds=None, samples=SAMPLES,
        workers=WORKERS, batch_size=BATCH_SIZE,
        verbose=True, sparse=True, *,
        precision='exact', device=None, return_info=False):
    """Mesh ``sdf`` (core.py:84-150).

    Returns the triangle soup as an ``(3*T, 3)`` float32 ndarray of world-space
    vertices, three consecutive rows per triangle -- ``len(points)``,
    ``points[:3]`` and ``write_binary_stl(path, points)`` behave as with the
    reference's list of float64 rows (README "generate").  Batch order is not
    preserved (triangles are emitted by whichever SM finishes first).
    """
    start = time.time()
    bounds, (dx, dy, dz), (X, Y, Z) = make_lattice(sdf, step, bounds, samples)
    (x0, y0, z0), (x1, y1, z1) = bounds

    if verbose:
        print('min %g, %g, %g' % (x0, y0, z0))
        print('max %g, %g, %g' % (x1, y1, z1))
        print('step %g, %g, %g' % (dx, dy, dz))

    num_batches = num_samples = None
    if verbose or return_info:
        (nbx, nby, nbz), num_samples = batch_counts(X, Y, Z, batch_size)
        num_batches = nbx * nby * nbz
    if verbose:
        print('%d samples in %d batches with %d workers' % (num_samples, num_batches, workers))

    bar = progress.Bar(num_batches or 0, enabled=verbose)
    try:
        tris, info = run_batches(sdf, X, Y, Z, batch_size, sparse, precision, device=device)
    except NotLowerableError:
        tris, info = _generate_opaque(sdf, X, Y, Z, batch_size, sparse, device)
    bar.done()
    points = tris.reshape(-1, 3)

    if verbose:
        print('%d skipped, %d empty, %d nonempty' % (info['skipped'], info['empty'], info['nonempty']))
        seconds = time.time() - start
        print('%d triangles in %g seconds' % (len(points) // 3, seconds))

    if return_info:
        info.update(batches=num_batches, samples=num_samples, dims=(len(X), len(Y), len(Z)))
        return points, info
    return points


def save_stl(path, sdf, step=None, bounds=None, samples=SAMPLES, workers=WORKERS, batch_size=BATCH_SIZE,
             verbose=True, sparse=True, *, precision='exact', device=None):
    """generate() + write_binary_stl() (core.py:152-156, stl.py:4-24) with the 50-byte records built on the
    device: the file payload comes back in one page-locked copy and is written as is."""
    import struct
    start = time.time()
    bounds, (dx, dy, dz), (X, Y, Z) = make_lattice(sdf, step, bounds, samples)
    (x0, y0, z0), (x1, y1, z1) = bounds
    if verbose:
        print('min %g, %g, %g' % (x0, y0, z0))
        print('max %g, %g, %g' % (x1, y1, z1))
        print('step %g, %g, %g' % (dx, dy, dz))
    (nbx, nby, nbz), num_samples = batch_counts(X, Y, Z, batch_size)
    if verbose:
        print('%d samples in %d batches with %d workers' % (num_samples, nbx * nby * nbz, workers))
    try:
        h, key = program_handle(sdf)
    except NotLowerableError:           # a Python closure in the tree: batch-by-batch route, records built on the host
        points = generate(sdf, step, bounds, samples, workers, batch_size, verbose, sparse, precision=precision, device=device)
        stl.write_binary_stl(path, points)
        return len(points) // 3
    X = np.ascontiguousarray(X, np.float64)
    Y = np.ascontiguousarray(Y, np.float64)
    Z = np.ascontiguousarray(Z, np.float64)
    tau = mixed_tau(h.program, X, Y, Z)
    dev = default_device() if device is None else device
    hint_key = (key, len(X), len(Y), len(Z), batch_size, bool(sparse), (0, -1))
    cap = _capacity_hint.get(hint_key, default_capacity(X, Y, Z))
    bar = progress.Bar(nbx * nby * nbz, enabled=verbose)
    while True:
        rec = _abi.pinned_pool.take((cap * 50 + 3) // 4)
        try:
            n_tri, stats, _ = h.generate_stl(X, Y, Z, batch_size, sparse, precision, tau, rec, cap, device=dev)
            break
        except _abi.CapacityError as e:
            cap = int(e.needed * 1.05) + 1024
    _capacity_hint[hint_key] = int(n_tri * 1.05) + 1024
    bar.done()
    with open(path, 'wb') as fp:
        fp.write(b'\x00' * 80)
        fp.write(struct.pack('<I', n_tri))
        fp.write(memoryview(rec).cast('B')[:n_tri * 50])
    if verbose:
        print('%d skipped, %d empty, %d nonempty' % (int(stats[0]), int(stats[1]), int(stats[2])))
        print('%d triangles in %g seconds' % (n_tri, time.time() - start))
    return n_tri


def save(path, *args, **kwargs):
    """core.py:152-158."""
    if path.lower().endswith('.stl'):
        save_stl(path, *args, **kwargs)
        return
    points = generate(*args, **kwargs)
    mesh = _mesh(points)
    mesh.write(path)


def weld(points, device=None):
    """The welding step of core.py:160-164, ``np.unique(points, axis=0, return_inverse=True)``, on the GPU:
    (unique vertices (V, 3) in lexicographic order, faces (T, 3) int32 indexing them)."""
    verts, inverse = _abi.weld(points, device=default_device() if device is None else device)
    return verts, inverse.reshape(-1, 3)


def _mesh(points):
    """core.py:160-164: the soup welded into an indexed mesh (on the GPU) as a ``meshio.Mesh`` -- or, where the
    optional meshio package is missing, as the built-in stand-in that writes .obj / .ply / .off (mesh_io.py)."""
    verts, faces = weld(points)
    try:
        import meshio
    except ImportError:
        from . import mesh_io
        return mesh_io.IndexedMesh(verts, [('triangle', faces)])
    return meshio.Mesh(verts, [('triangle', faces)])


def sample_slice(sdf, w=1024, h=1024, x=None, y=None, z=None, bounds=None):
    """core.py:202-232: evaluate one axis-aligned w x h slice (on the GPU, float64)."""
    if bounds is None:
        bounds = _estimate_bounds(sdf)
    (x0, y0, z0), (x1, y1, z1) = bounds
    if x is not None:
        axes, names = (np.array([x]), np.linspace(y0, y1, w), np.linspace(z0, z1, h)), 'ZY'
        extent = (axes[2][0], axes[2][-1], axes[1][0], axes[1][-1])
    elif y is not None:
        axes, names = (np.linspace(x0, x1, w), np.array([y]), np.linspace(z0, z1, h)), 'ZX'
        extent = (axes[2][0], axes[2][-1], axes[0][0], axes[0][-1])
    elif z is not None:
        axes, names = (np.linspace(x0, x1, w), np.linspace(y0, y1, h), np.array([z])), 'YX'
        extent = (axes[1][0], axes[1][-1], axes[0][0], axes[0][-1])
    else:
        raise Exception('x, y, or z position must be specified')
    P = _cartesian_product(*axes)
    return sdf(P).reshape((w, h)), extent, names


def show_slice(*args, **kwargs):
    """core.py:234-244 (needs matplotlib, as in the reference)."""
    import matplotlib.pyplot as plt
    show_abs = kwargs.pop('abs', False)
    a, extent, axes = sample_slice(*args, **kwargs)
    if show_abs:
        a = np.abs(a)
    im = plt.imshow(a, extent=extent, origin='lower')
    plt.xlabel(axes[0])
    plt.ylabel(axes[1])
    plt.colorbar(im)
    plt.show()

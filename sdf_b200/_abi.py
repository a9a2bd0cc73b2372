"""ctypes binding of libsdfb200.so (include/sdfb200.h).

This is the whole boundary between the Python host code and the CUDA path; it
is exactly the stub a maintainer of the reference would add.
There is no CPU fallback: a missing library or a missing GPU raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libsdfb200.so')

OK, E_INVALID, E_CUDA, E_CAPACITY, E_NOGPU = 0, -1, -2, -3, -4
PREC = {'f32': 0, 'mixed': 1, 'f64': 2, 'exact': 3}

EXPORTS = ['sdfb200_abi_version', 'sdfb200_last_error', 'sdfb200_device_count',
           'sdfb200_program_create', 'sdfb200_program_free', 'sdfb200_eval_f32', 'sdfb200_eval_f32_fast',
           'sdfb200_eval_f64', 'sdfb200_generate', 'sdfb200_skip_flags',
           'sdfb200_host_alloc', 'sdfb200_host_free', 'sdfb200_host_register', 'sdfb200_host_unregister', 'sdfb200_copy_to_host', 'sdfb200_jit_probe', 'sdfb200_jit_plan', 'sdfb200_generate_stl',
           'sdfb200_shared_create', 'sdfb200_shared_open', 'sdfb200_shared_free', 'sdfb200_shared_buffer',
           'sdfb200_shared_wait', 'sdfb200_shared_release', 'sdfb200_generate_shared', 'sdfb200_generate_window', 'sdfb200_mc_volume', 'sdfb200_weld']


class SdfB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__('libsdfb200 error %d: %s' % (code, message))
        self.code = code


class CapacityError(SdfB200Error):
    def __init__(self, code, message, needed):
        super().__init__(code, message)
        self.needed = needed


_lib = None


def lib():
    """Load the shared library (built in-tree by ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError('%s is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                          '(sdf_b200 has no CPU fallback)' % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double
    L.sdfb200_abi_version.restype = i32
    L.sdfb200_last_error.restype = ctypes.c_char_p
    L.sdfb200_device_count.argtypes = [ctypes.POINTER(i32)]
    L.sdfb200_program_create.argtypes = [vp, i32, vp, i32, ctypes.POINTER(vp)]
    L.sdfb200_program_free.argtypes = [vp]
    L.sdfb200_program_free.restype = None
    L.sdfb200_eval_f32.argtypes = [vp, vp, i64, vp, i32, vp]
    L.sdfb200_eval_f64.argtypes = [vp, vp, i64, vp, i32, vp]
    L.sdfb200_eval_f32_fast.argtypes = [vp, vp, i64, vp, i32, vp]
    L.sdfb200_generate.argtypes = [vp, vp, i64, vp, i64, vp, i64, i32, i32, i32, dbl, i32, i32,
                                   vp, i64, ctypes.POINTER(i64), vp, vp, i32, vp]
    L.sdfb200_host_alloc.argtypes = [ctypes.c_size_t, ctypes.POINTER(vp)]
    L.sdfb200_host_free.argtypes = [vp]
    L.sdfb200_host_free.restype = None
    L.sdfb200_host_register.argtypes = [vp, ctypes.c_size_t]
    L.sdfb200_host_unregister.argtypes = [vp]
    L.sdfb200_host_unregister.restype = None
    L.sdfb200_copy_to_host.argtypes = [vp, vp, ctypes.c_size_t, i32, vp]
    L.sdfb200_generate_stl.argtypes = [vp, vp, i64, vp, i64, vp, i64, i32, i32, i32, dbl,
                                       vp, i64, ctypes.POINTER(i64), vp, vp, i32, vp]
    L.sdfb200_shared_create.argtypes = [i64, i32, ctypes.POINTER(vp), vp]
    L.sdfb200_shared_open.argtypes = [vp, i64, i32, ctypes.POINTER(vp)]
    L.sdfb200_shared_free.argtypes = [vp]
    L.sdfb200_shared_free.restype = None
    L.sdfb200_shared_buffer.argtypes = [vp, i32]
    L.sdfb200_shared_buffer.restype = vp
    L.sdfb200_shared_wait.argtypes = [vp, i32, i32, ctypes.POINTER(i64), vp]
    L.sdfb200_shared_release.argtypes = [vp, i32, vp]
    L.sdfb200_generate_shared.argtypes = [vp, vp, i64, vp, i64, vp, i64, i32, i32, i32, dbl, i32, i32,
                                          vp, i32, ctypes.POINTER(i64), vp, vp, i32, vp]
    L.sdfb200_generate_window.argtypes = [vp, vp, i64, vp, i64, vp, i64, i32, i32, i32, dbl, i64, i64,
                                          vp, i64, vp, i32, ctypes.POINTER(i64), vp, vp, i32, vp]
    L.sdfb200_mc_volume.argtypes = [vp, vp, i64, i64, i64, vp, vp, vp, vp, i64, ctypes.POINTER(i64), i32, vp]
    L.sdfb200_weld.argtypes = [vp, i64, vp, vp, ctypes.POINTER(i64), i32, vp]
    L.sdfb200_jit_probe.argtypes = [vp, i32, ctypes.POINTER(i64)]
    L.sdfb200_jit_plan.argtypes = [vp, i32, ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(dbl)]
    L.sdfb200_skip_flags.argtypes = [vp, vp, i64, vp, i64, vp, i64, i32, vp, i32, vp]
    _lib = L
    return L


def _check(rc):
    if rc != OK:
        msg = lib().sdfb200_last_error().decode('utf-8', 'replace')
        raise SdfB200Error(rc, msg)


def device_count():
    n = ctypes.c_int(0)
    _check(lib().sdfb200_device_count(ctypes.byref(n)))
    return n.value


def _ptr(obj):
    """Address of a numpy array or a torch tensor (host or device)."""
    if obj is None:
        return None
    if isinstance(obj, np.ndarray):
        return obj.ctypes.data
    return obj.data_ptr()


def host_register(address, nbytes):
    _check(lib().sdfb200_host_register(address, nbytes))


def host_unregister(address):
    if _lib is not None:
        _lib.sdfb200_host_unregister(address)


def copy_to_host(dst_address, src_device_ptr, nbytes, device=0, stream=None):
    _check(lib().sdfb200_copy_to_host(dst_address, src_device_ptr, nbytes, device, stream))


def weld(points, device=0, stream=None):
    """np.unique(points, axis=0, return_inverse=True) of a (3T, 3) float32 vertex soup on the GPU
    (include/sdfb200.h sdfb200_weld): (unique vertices (V, 3) float32 in lexicographic order, inverse (3T,) int32)."""
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    if len(pts) % 3:
        raise ValueError('a triangle soup has 3 vertices per triangle')
    n_tri = len(pts) // 3
    verts = np.empty((max(len(pts), 1), 3), np.float32)
    faces = np.empty(max(len(pts), 1), np.int32)
    nv = ctypes.c_int64(0)
    _check(lib().sdfb200_weld(pts.ctypes.data, n_tri, verts.ctypes.data, faces.ctypes.data, ctypes.byref(nv), device, stream))
    return verts[:nv.value], faces[:len(pts)]


class PinnedPool:
    """Page-locked output buffers, recycled when the arrays handed to the user are garbage collected.

    ``take(n_floats)`` returns a float32 ndarray over pinned memory; every view of it keeps the
    block alive, and the block goes back to the pool (it is not freed) once the last view dies --
    so results stay valid for as long as the caller holds them, and a loop that drops its previous
    result re-uses the same pinned block without any allocation or extra copy.
    """

    def __init__(self, max_free=4):
        self.free = []          # (nbytes, address)
        self.max_free = max_free

    def take(self, n_floats):
        import weakref
        nbytes = max(int(n_floats) * 4, 4096)
        best = None
        for i, (sz, addr) in enumerate(self.free):
            if sz >= nbytes and (best is None or sz < self.free[best][0]):
                best = i
        if best is not None:
            sz, addr = self.free.pop(best)
        else:
            sz = 1 << max(12, (nbytes - 1).bit_length())          # round up: fewer distinct sizes
            p = ctypes.c_void_p()
            _check(lib().sdfb200_host_alloc(sz, ctypes.byref(p)))
            addr = p.value
        base = np.ctypeslib.as_array((ctypes.c_float * (sz // 4)).from_address(addr))
        weakref.finalize(base, self._release, sz, addr)
        return base

    def _release(self, sz, addr):
        if len(self.free) < self.max_free:
            self.free.append((sz, addr))
        elif _lib is not None:
            _lib.sdfb200_host_free(addr)


pinned_pool = PinnedPool()


class SharedOutput:
    """Two triangle buffers (step parity) + flag words in the root GPU's memory that every rank's extraction kernel
    writes into (include/sdfb200.h "fused multi-GPU gather").  ``handle`` (64 bytes) is what the root broadcasts."""

    def __init__(self, capacity, device, handle=None):
        self._h = ctypes.c_void_p()
        self.capacity, self.device, self.owner = int(capacity), int(device), handle is None
        if handle is None:
            buf = (ctypes.c_ubyte * 64)()
            _check(lib().sdfb200_shared_create(self.capacity, self.device, ctypes.byref(self._h), buf))
            self.handle = bytes(buf)
        else:
            self.handle = bytes(handle)
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(self.handle)
            _check(lib().sdfb200_shared_open(buf, self.capacity, self.device, ctypes.byref(self._h)))

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h and _lib is not None:
            _lib.sdfb200_shared_free(h)

    def buffer_ptr(self, which):
        return lib().sdfb200_shared_buffer(self._h, int(which))

    def wait(self, which, world, stream=None):
        """Root only: device-side wait until all ``world`` ranks have delivered the step of parity ``which``;
        returns the triangle count (synchronises ``stream``)."""
        n = ctypes.c_int64(0)
        _check(lib().sdfb200_shared_wait(self._h, int(which), int(world), ctypes.byref(n), stream))
        return n.value

    def release(self, which, stream=None):
        """Root only: the buffer of parity ``which`` has been consumed; hands it to the step after next."""
        _check(lib().sdfb200_shared_release(self._h, int(which), stream))


class ProgramHandle:
    """Owns one sdfb200_prog (freed on garbage collection)."""

    def __init__(self, program):
        self._h = ctypes.c_void_p()
        self.program = program
        code = np.ascontiguousarray(program.code, np.uint32)
        consts = np.ascontiguousarray(program.consts, np.float64)
        _check(lib().sdfb200_program_create(code.ctypes.data, len(code),
                                            consts.ctypes.data if len(consts) else None, len(consts),
                                            ctypes.byref(self._h)))

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h and _lib is not None:
            _lib.sdfb200_program_free(h)

    def eval_f32(self, xyz, out, n, device=0, stream=None):
        _check(lib().sdfb200_eval_f32(self._h, _ptr(xyz), n, _ptr(out), device, stream))

    def jit_probe(self, kernel):
        """NVRTC-compile the specialised kernel `kernel` (no GPU needed); returns the cubin size."""
        n = ctypes.c_int64(0)
        _check(lib().sdfb200_jit_probe(self._h, kernel, ctypes.byref(n)))
        return n.value

    def jit_plan(self, precision=3):
        """(share axis, samples per thread, invariant fraction) of the specialised tile kernel (no GPU needed)."""
        ax, n, fr = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_double(0)
        _check(lib().sdfb200_jit_plan(self._h, precision, ctypes.byref(ax), ctypes.byref(n), ctypes.byref(fr)))
        return ax.value, n.value, fr.value

    def eval_f32_fast(self, xyz, out, n, device=0, stream=None):
        _check(lib().sdfb200_eval_f32_fast(self._h, _ptr(xyz), n, _ptr(out), device, stream))

    def eval_f64(self, xyz, out, n, device=0, stream=None):
        _check(lib().sdfb200_eval_f64(self._h, _ptr(xyz), n, _ptr(out), device, stream))

    def skip_flags(self, X, Y, Z, batch_size, device=0, stream=None):
        nb = [(len(a) + batch_size - 1) // batch_size for a in (X, Y, Z)]
        flags = np.zeros(nb[0] * nb[1] * nb[2], np.uint8)
        _check(lib().sdfb200_skip_flags(self._h, X.ctypes.data, len(X), Y.ctypes.data, len(Y),
                                        Z.ctypes.data, len(Z), batch_size, flags.ctypes.data, device, stream))
        return flags.reshape(nb)

    def generate_stl(self, X, Y, Z, batch_size, sparse, precision, tau, records_out, capacity, device=0, stream=None):
        """Like generate(), but fills 50-byte binary-STL records; returns (n_tri, stats[6], timings_ms[4])."""
        n_tri = ctypes.c_int64(0)
        stats = np.zeros(6, np.int64)
        timings = np.zeros(8, np.float64)
        rc = lib().sdfb200_generate_stl(self._h, X.ctypes.data, len(X), Y.ctypes.data, len(Y), Z.ctypes.data, len(Z),
                                        int(batch_size), int(bool(sparse)), PREC[precision], float(tau),
                                        _ptr(records_out), int(capacity), ctypes.byref(n_tri), stats.ctypes.data,
                                        timings.ctypes.data, device, stream)
        if rc == E_CAPACITY:
            raise CapacityError(rc, lib().sdfb200_last_error().decode(), n_tri.value)
        _check(rc)
        return n_tri.value, stats, timings

    def generate_shared(self, X, Y, Z, batch_size, sparse, precision, tau, shared, which, zb=(0, -1), device=0,
                        stream=None):
        """This rank's slab, written straight into the root's shared buffer; returns (n_local, stats, timings)."""
        n_tri = ctypes.c_int64(0)
        stats = np.zeros(6, np.int64)
        timings = np.zeros(8, np.float64)
        _check(lib().sdfb200_generate_shared(self._h, X.ctypes.data, len(X), Y.ctypes.data, len(Y), Z.ctypes.data, len(Z),
                                             int(batch_size), int(bool(sparse)), PREC[precision], float(tau),
                                             int(zb[0]), int(zb[1]), shared._h, int(which), ctypes.byref(n_tri),
                                             stats.ctypes.data, timings.ctypes.data, device, stream))
        return n_tri.value, stats, timings

    def mc_volume(self, volume, X, Y, Z, device=0, stream=None):
        """Marching cubes + world mapping of one batch volume computed by the caller (include/sdfb200.h
        sdfb200_mc_volume); returns a (T, 9) float32 ndarray."""
        vol = np.ascontiguousarray(volume, np.float32)
        X, Y, Z = (np.ascontiguousarray(a, np.float64) for a in (X, Y, Z))
        assert vol.shape == (len(X), len(Y), len(Z)), (vol.shape, len(X), len(Y), len(Z))
        cap = max(1, (len(X) - 1) * (len(Y) - 1) * (len(Z) - 1)) * 5
        n_tri = ctypes.c_int64(0)
        while True:
            out = np.empty((cap, 9), np.float32)
            rc = lib().sdfb200_mc_volume(self._h, vol.ctypes.data, len(X), len(Y), len(Z), X.ctypes.data, Y.ctypes.data,
                                         Z.ctypes.data, out.ctypes.data, cap, ctypes.byref(n_tri), device, stream)
            if rc == E_CAPACITY:
                cap = int(n_tri.value) + 16
                continue
            _check(rc)
            return out[:n_tri.value]

    def generate_window(self, X, Y, Z, batch_size, sparse, precision, tau, window, tri_out=None, capacity=0,
                        shared=None, which=0, device=0, stream=None):
        """The batches [window[0], window[1]) of the Z-major batch order (include/sdfb200.h); output into
        ``tri_out`` (device / host / None = count only) or, with ``shared``, into the root GPU's buffer.
        Returns (n_tri, stats[6], timings_ms[7]); raises CapacityError with .needed."""
        n_tri = ctypes.c_int64(0)
        stats = np.zeros(6, np.int64)
        timings = np.zeros(8, np.float64)
        rc = lib().sdfb200_generate_window(self._h, X.ctypes.data, len(X), Y.ctypes.data, len(Y), Z.ctypes.data, len(Z),
                                           int(batch_size), int(bool(sparse)), PREC[precision], float(tau),
                                           int(window[0]), int(window[1]), _ptr(tri_out), int(capacity),
                                           shared._h if shared is not None else None, int(which),
                                           ctypes.byref(n_tri), stats.ctypes.data, timings.ctypes.data, device, stream)
        if rc == E_CAPACITY:
            raise CapacityError(rc, lib().sdfb200_last_error().decode(), n_tri.value)
        _check(rc)
        return n_tri.value, stats, timings

    def generate(self, X, Y, Z, batch_size, sparse, precision, tau, tri_out, capacity,
                 zb=(0, -1), device=0, stream=None):
        """Returns (n_tri, stats[6], timings_ms[4]); raises CapacityError with .needed."""
        n_tri = ctypes.c_int64(0)
        stats = np.zeros(6, np.int64)
        timings = np.zeros(8, np.float64)
        rc = lib().sdfb200_generate(self._h, X.ctypes.data, len(X), Y.ctypes.data, len(Y),
                                    Z.ctypes.data, len(Z), int(batch_size), int(bool(sparse)),
                                    PREC[precision], float(tau), int(zb[0]), int(zb[1]), _ptr(tri_out), int(capacity),
                                    ctypes.byref(n_tri), stats.ctypes.data, timings.ctypes.data,
                                    device, stream)
        if rc == E_CAPACITY:
            raise CapacityError(rc, lib().sdfb200_last_error().decode(), n_tri.value)
        _check(rc)
        return n_tri.value, stats, timings

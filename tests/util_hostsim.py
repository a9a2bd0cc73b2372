"""Builds and loads the host execution of the device code (tests/hostsim/): TEST INFRASTRUCTURE.

The kernel bodies in sdf_b200/csrc/{sdf_ops.cuh, sdf_kernels.cuh} are compiled by g++ against
tests/hostsim/cuda_runtime.h (CUDA threads as fibers) from PATCHED COPIES made here -- the product files are
not touched.  The patches replace what no host compiler can take: five inline-PTX helpers (mbarrier, TMA bulk
copy, an acquire load, the approximate square root) and the one `extern __shared__` declaration.  Every patch
asserts that it matched exactly as often as expected, so a change of the device source that the stand-in does
not cover fails loudly instead of being skipped.
"""
import ctypes
import hashlib
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'sdf_b200', 'csrc')
HS = os.path.join(ROOT, 'tests', 'hostsim')
BUILD = os.path.join(HS, '_build')

PREC = {'f32': 0, 'mixed': 1, 'f64': 2, 'exact': 3}


def _sub(text, pattern, repl, count, flags=0):
    out, n = re.subn(pattern, repl, text, flags=flags)
    assert n == count, 'hostsim patch %r matched %d times, expected %d: the device source changed' % (pattern, n, count)
    return out


def _body(name, new_body):
    """(pattern, replacement) that swaps the body of the top-level device function `name`."""
    pat = r'(__device__ __forceinline__ [^\n;{]*\b%s\([^)]*\) \{\n).*?\n\}\n' % re.escape(name)
    return pat, lambda m: m.group(1) + '    ' + new_body + '\n}\n'


def patched_sources(dst):
    os.makedirs(dst, exist_ok=True)
    for f in ('sdf_opcodes.h', 'mc_tables.inc'):
        shutil.copy(os.path.join(CSRC, f), os.path.join(dst, f))
    ops = open(os.path.join(CSRC, 'sdf_ops.cuh')).read()
    ops = _sub(ops, r'asm\("sqrt\.approx\.ftz\.f32 %0, %1;" : "=f"\(r\) : "f"\(x\)\);', 'r = hostsim::sqrt_approx(x);', 1)
    assert 'asm' not in re.sub(r'//[^\n]*', '', ops), 'unpatched inline assembly in sdf_ops.cuh'
    open(os.path.join(dst, 'sdf_ops.cuh'), 'w').write(ops)
    k = open(os.path.join(CSRC, 'sdf_kernels.cuh')).read()
    for name, body in (('mbar_init', 'hostsim::mbar_init(bar, count);'),
                       ('mbar_expect_tx', 'hostsim::mbar_expect_tx(bar, bytes);'),
                       ('mbar_wait', 'hostsim::mbar_wait(bar, parity);'),
                       ('tma_bulk_g2s', 'hostsim::tma_bulk_g2s(dst_smem, src_gmem, bytes, bar);'),
                       ('ld_acquire_sys_u64', 'return *p;')):
        pat, repl = _body(name, body)
        k = _sub(k, pat, repl, 1, flags=re.S)
    k = _sub(k, r'asm volatile\("fence\.mbarrier_init\.release\.cluster;" ::: "memory"\);', ';', 2)
    k = _sub(k, r'extern __shared__ __align__\(16\) unsigned char smem_raw\[\];',
             'unsigned char* const smem_raw = hostsim::dyn_smem();', 1)
    assert 'asm' not in re.sub(r'//[^\n]*', '', k), 'unpatched inline assembly in sdf_kernels.cuh'
    # every __shared__ variable is announced, so that each CTA finds garbage in it instead of its predecessor's data
    k = _sub(k, r'(\n[ \t]+__shared__ [^;\n]*?\b(\w+)(?:\[[^;\n]*\])?;)', r'\1 hostsim::note_shared(&\2, sizeof(\2));', 6)
    open(os.path.join(dst, 'sdf_kernels.cuh'), 'w').write(k)


def rewrite_launches(text):
    """`kernel<targs><<<grid, block, smem, stream>>>(args);` -> `hostsim::launch_kernel(kernel<targs>, grid, block, smem,
    stream, args);` (the one piece of CUDA syntax a host compiler cannot parse).  Returns (text, launches rewritten)."""
    out, pos, count = [], 0, 0
    while True:
        k = text.find('<<<', pos)
        if k < 0:
            break
        # the kernel expression: identifier, optionally followed by a template argument list
        b = k
        if text[b - 1] == '>':
            depth = 0
            while True:
                b -= 1
                depth += {'>': 1, '<': -1}.get(text[b], 0)
                if depth == 0:
                    break
        while b > 0 and (text[b - 1].isalnum() or text[b - 1] in '_:'):
            b -= 1
        e = text.index('>>>', k)
        assert text[e + 3] == '(', text[k - 40:e + 10]
        depth, a = 0, e + 3
        while True:
            depth += {'(': 1, ')': -1}.get(text[a], 0)
            if depth == 0:
                break
            a += 1
        args = text[e + 4:a].strip()
        out.append(text[pos:b])
        out.append('hostsim::launch_kernel(%s, %s%s)' % (text[b:k], text[k + 3:e], ', ' + args if args else ''))
        pos = a + 1
        count += 1
    out.append(text[pos:])
    return ''.join(out), count


def patched_library_sources(dst):
    """sdfb200.cu / sdfb200_weld.cu for the host compiler: launches rewritten, nothing else touched."""
    patched_sources(dst)
    shutil.copy(os.path.join(CSRC, 'sdf_jit_entry.cuh'), os.path.join(dst, 'sdf_jit_entry.cuh'))
    expected = {'sdfb200.cu': 20, 'sdfb200_weld.cu': 4}
    for f, n_expected in expected.items():
        text = open(os.path.join(CSRC, f)).read()
        text, n = rewrite_launches(text)
        assert n == n_expected, '%s: %d launches rewritten, expected %d: the host source changed' % (f, n, n_expected)
        text = text.replace('#include "../../include/sdfb200.h"', '#include "%s"' % os.path.join(ROOT, 'include', 'sdfb200.h'))
        if f == 'sdfb200.cu':
            # two sizing constants become knobs of the host build, so that a 64^3 lattice can drive the paths a GPU only
            # takes at sizes no emulation reaches: rounds cut by the HOST (tile volumes beyond the 24 GB budget) and the
            # retry after a work-list overflow ($HOSTSIM_VOL_BUDGET bytes, $HOSTSIM_MIN_CUBES entries)
            text = _sub(text, r'static const size_t kVolBudgetBytes = \(size_t\)24 << 30;',
                        '#define kVolBudgetBytes (getenv("HOSTSIM_VOL_BUDGET") ? (size_t)atoll(getenv("HOSTSIM_VOL_BUDGET")) '
                        ': (size_t)24 << 30)', 1)
            # $HOSTSIM_NVRTC: the stand-in compiler (tests/hostsim/fake_nvrtc.cpp) instead of libnvrtc
            text = _sub(text, r'const char\* names\[\] = \{"libnvrtc\.so\.12",',
                        'const char* names[] = {getenv("HOSTSIM_NVRTC") ? getenv("HOSTSIM_NVRTC") : "libnvrtc.so.12",', 1)
            text = _sub(text, r'std::max<size_t>\(\(size_t\)1 << 20, est_tiles \* 3072\)',
                        'std::max<size_t>(getenv("HOSTSIM_MIN_CUBES") ? (size_t)atoll(getenv("HOSTSIM_MIN_CUBES")) : (size_t)1 << 20, '
                        '(getenv("HOSTSIM_MIN_CUBES") ? (size_t)0 : est_tiles * 3072))', 1)
        open(os.path.join(dst, f.replace('.cu', '.cpp')), 'w').write(text)


def build_fake_nvrtc():
    """tests/hostsim/fake_nvrtc.cpp as a shared library; returns (path, directory of the patched device headers)."""
    srcs = [os.path.join(CSRC, f) for f in ('sdf_ops.cuh', 'sdf_kernels.cuh', 'sdf_opcodes.h', 'mc_tables.inc', 'sdf_jit_entry.cuh')]
    srcs += [os.path.join(HS, 'fake_nvrtc.cpp'), os.path.join(HS, 'cuda_runtime.h'), os.path.abspath(__file__)]
    key = _digest(*srcs)
    so = os.path.join(BUILD, 'libhostsim_nvrtc_%s.so' % key)
    inc = os.path.join(BUILD, 'jitinc_%s' % key)
    if not os.path.exists(so):
        patched_sources(inc)
        shutil.copy(os.path.join(CSRC, 'sdf_jit_entry.cuh'), os.path.join(inc, 'sdf_jit_entry.cuh'))
        shutil.copy(os.path.join(HS, 'cuda_runtime.h'), os.path.join(inc, 'cuda_runtime.h'))      # a snapshot: units compiled
        tmp = so + '.%d.tmp' % os.getpid()                                                         # later must match this library
        r = subprocess.run(['g++', '-std=c++17', '-O1', '-fPIC', '-shared', os.path.join(HS, 'fake_nvrtc.cpp'), '-o', tmp],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-4000:]
        os.replace(tmp, so)
    return so, inc


def build_library(sanitize=None):
    """``sanitize`` (default: $HOSTSIM_SANITIZE): ASan + UBSan instrumentation; the process must preload libasan.
    The WHOLE C-ABI library (sdfb200.cu + sdfb200_weld.cu + the kernels) compiled for the host: same exports as
    sdf_b200/libsdfb200.so, kernels executed by the SIMT stand-in, CUDA runtime calls by its host API stand-in."""
    if shutil.which('g++') is None:
        pytest.skip('no g++')
    srcs = [os.path.join(CSRC, f) for f in ('sdf_ops.cuh', 'sdf_kernels.cuh', 'sdf_opcodes.h', 'mc_tables.inc', 'sdfb200.cu',
                                            'sdfb200_weld.cu', 'sdf_jit_entry.cuh')]
    srcs += [os.path.join(HS, 'cuda_runtime.h'), os.path.join(HS, 'cub', 'device', 'device_radix_sort.cuh'),
             os.path.join(HS, 'cub', 'device', 'device_scan.cuh'), os.path.join(ROOT, 'include', 'sdfb200.h'),
             os.path.abspath(__file__)]
    if sanitize is None:
        sanitize = bool(os.environ.get('HOSTSIM_SANITIZE'))
    coverage = bool(os.environ.get('HOSTSIM_COVERAGE'))         # gcov instrumentation (tools/hostsim_coverage.py --library)
    key = _digest(*srcs, extra=('asan' if sanitize else '') + ('gcov' if coverage else ''))
    so = os.path.join(BUILD, 'libsdfb200_hostsim_%s.so' % key)
    if os.path.exists(so):
        return so
    work = os.path.join(BUILD, 'lib_%s' % key)
    patched_library_sources(work)
    flags = ['-std=c++17', '-O1', '-fPIC', '-shared', '-ffp-contract=off', '-fno-strict-aliasing', '-Wno-unknown-pragmas',
             '-I', work, '-I', HS]
    if sanitize:
        flags += ['-g', '-fsanitize=address,undefined', '-fno-sanitize-recover=undefined', '-fno-omit-frame-pointer']
    if coverage:
        flags = [f for f in flags if f != '-O1'] + ['-O0', '-g', '--coverage']
    tmp = so + '.%d.tmp' % os.getpid()
    r = subprocess.run(['g++'] + flags + [os.path.join(work, 'sdfb200.cpp'), os.path.join(work, 'sdfb200_weld.cpp'), '-o', tmp,
                                          '-ldl', '-lpthread'], capture_output=True, text=True, cwd=work)
    assert r.returncode == 0, 'hostsim library build failed:\n' + r.stderr[-8000:]
    os.replace(tmp, so)
    return so


def _digest(*paths, extra=''):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        h.update(open(p, 'rb').read())
    return h.hexdigest()[:16]


def sanitizer_preload():
    """LD_PRELOAD value for a process that loads a harness built with ``sanitize=True`` (None when gcc has no libasan)."""
    libs = []
    for name in ('libasan.so', 'libubsan.so'):
        path = subprocess.run(['g++', '-print-file-name=' + name], capture_output=True, text=True).stdout.strip()
        if not os.path.isabs(path) or not os.path.exists(path):
            return None
        libs.append(path)
    return ':'.join(libs)


def build(spec_text=None, sanitize=False):
    """The harness as a shared library; with ``spec_text`` (a `struct SpecEval {...};`) the specialised evaluator;
    ``sanitize``: AddressSanitizer + UndefinedBehaviorSanitizer instrumentation (the memcheck of this execution)."""
    if shutil.which('g++') is None:
        pytest.skip('no g++')
    srcs = [os.path.join(CSRC, f) for f in ('sdf_ops.cuh', 'sdf_kernels.cuh', 'sdf_opcodes.h', 'mc_tables.inc')]
    srcs += [os.path.join(HS, 'cuda_runtime.h'), os.path.join(HS, 'harness.cpp'), os.path.abspath(__file__)]
    key = _digest(*srcs, extra=(spec_text or '') + ('|asan' if sanitize else ''))
    so = os.path.join(BUILD, 'hostsim_%s.so' % key)
    if os.path.exists(so):
        return so
    work = os.path.join(BUILD, 'src_%s' % key)
    patched_sources(work)
    flags = ['-std=c++17', '-O1', '-fPIC', '-shared', '-ffp-contract=off', '-fno-strict-aliasing', '-Wno-unknown-pragmas',
             '-I', work, '-I', HS]
    if spec_text is not None:
        open(os.path.join(work, 'spec_eval.inc'), 'w').write(spec_text)
        flags.append('-DHOSTSIM_SPEC="spec_eval.inc"')
    if sanitize:
        flags += ['-g', '-fsanitize=address,undefined', '-fno-sanitize-recover=undefined', '-fno-omit-frame-pointer']
    tmp = so + '.%d.tmp' % os.getpid()
    r = subprocess.run(['g++'] + flags + [os.path.join(HS, 'harness.cpp'), '-o', tmp], capture_output=True, text=True)
    assert r.returncode == 0, 'hostsim build failed:\n' + r.stderr[-6000:]
    os.replace(tmp, so)
    return so


def spec_eval_text(program, kernel=4, env=None):
    """`struct SpecEval` exactly as sdfb200.cu's jit_source() generates it for this program (the library writes the
    source of a specialised kernel to $SDFB200_JIT_DUMP; NVRTC compiles without a GPU)."""
    import tempfile
    from sdf_b200 import _abi
    d = tempfile.mkdtemp(prefix='hostsim_jit')
    old = {k: os.environ.get(k) for k in ('SDFB200_JIT_DUMP', 'SDFB200_JIT_CACHE')}
    os.environ['SDFB200_JIT_DUMP'] = d
    os.environ['SDFB200_JIT_CACHE'] = os.path.join(d, 'cache')
    for k, v in (env or {}).items():
        old.setdefault(k, os.environ.get(k))
        os.environ[k] = str(v)
    try:
        h = _abi.ProgramHandle(program)
        h.jit_probe(kernel)
        src = open(os.path.join(d, 'jit_k%d.cu' % kernel)).read()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        shutil.rmtree(d, ignore_errors=True)
    m = re.search(r'struct SpecEval \{.*?\n\};\n', src, flags=re.S)
    assert m, 'no SpecEval in the generated source'
    return m.group(0)


class HostSim:
    """ctypes front end of the harness for one lowered program (``sdf_b200.ir.lower`` output)."""

    def __init__(self, program, spec_text=None, sanitize=False):
        self.lib = ctypes.CDLL(build(spec_text, sanitize))
        self.code = np.ascontiguousarray(program.code, np.uint32)
        self.consts = np.ascontiguousarray(program.consts, np.float64)
        if len(self.consts) == 0:
            self.consts = np.zeros(1, np.float64)
        self.n_consts = len(program.consts)
        self.lib.hs_set_fast_ulps.argtypes = [ctypes.c_double]
        self.lib.hs_switches.restype = ctypes.c_ulong

    def _prog(self):
        return (self.code.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(self.code)),
                self.consts.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(self.n_consts))

    def shape(self):
        kn, ks = ctypes.c_int(0), ctypes.c_int(0)
        spec = self.lib.hs_spec_shape(ctypes.byref(kn), ctypes.byref(ks))
        return bool(spec), kn.value, ks.value

    def set_fast_ulps(self, ulps):
        self.lib.hs_set_fast_ulps(float(ulps))

    def set_schedule(self, mode):
        """Order in which the fibers of a CTA run: 0 ascending, 1 descending, >= 2 a seeded random permutation per pass."""
        self.lib.hs_set_schedule(int(mode))

    def eval_f64(self, P, grid=3):
        P = np.ascontiguousarray(P, np.float64)
        out = np.empty(len(P), np.float64)
        rc = self.lib.hs_eval_f64(*self._prog(), P.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(len(P)),
                                  out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(grid))
        assert rc == 0, rc
        return out

    def eval_f32(self, P, fast=False, grid=2):
        P = np.ascontiguousarray(P, np.float32)
        out = np.empty(len(P), np.float32)
        rc = self.lib.hs_eval_f32(*self._prog(), P.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(len(P)),
                                  out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(int(fast)), ctypes.c_int(grid))
        assert rc == 0, rc
        return out

    def skip_flags(self, X, Y, Z, batch_size):
        X, Y, Z = (np.ascontiguousarray(a, np.float64) for a in (X, Y, Z))
        nb = [(len(a) + batch_size - 1) // batch_size for a in (X, Y, Z)]
        flags = np.zeros(nb[0] * nb[1] * nb[2], np.uint8)
        rc = self.lib.hs_skip_flags(*self._prog(), X.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(X)),
                                    Y.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(Y)),
                                    Z.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(Z)), ctypes.c_int(batch_size),
                                    flags.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0, rc
        return flags.reshape(nb)

    def generate(self, X, Y, Z, batch_size=32, sparse=True, precision='exact', tau=0.0, grids=(2, 3, 2, 3), cap=None):
        """(triangles (T, 9) float32, stats dict) of one sdfb200_generate step run on the host."""
        X, Y, Z = (np.ascontiguousarray(a, np.float64) for a in (X, Y, Z))
        g = np.asarray(grids, np.int32)
        n_tri = ctypes.c_longlong(0)
        stats = np.zeros(6, np.int64)

        def call(out, capacity):
            rc = self.lib.hs_generate(*self._prog(), X.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(X)),
                                      Y.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(Y)),
                                      Z.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(Z)), ctypes.c_int(batch_size),
                                      ctypes.c_int(int(bool(sparse))), ctypes.c_int(PREC[precision]), ctypes.c_double(tau),
                                      out.ctypes.data_as(ctypes.c_void_p) if out is not None else None,
                                      ctypes.c_longlong(capacity), ctypes.byref(n_tri),
                                      stats.ctypes.data_as(ctypes.c_void_p), g.ctypes.data_as(ctypes.c_void_p))
            assert rc == 0, rc
        if cap is None:
            call(None, 0)                                   # count pass, as the library's callers do
            cap = n_tri.value
        guard = 64                                          # rows past the capacity must stay untouched
        out = np.full((max(cap, 1) + guard, 9), np.float32(-12345.0))
        call(out, cap)
        assert np.all(out[max(cap, 1):] == np.float32(-12345.0)), 'the extraction kernel wrote past the capacity'
        keys = ('skipped', 'empty', 'nonempty', 'evaluated', 'refined', 'active_cubes')
        return out[:min(n_tri.value, cap)], dict(zip(keys, (int(v) for v in stats)), triangles=n_tri.value)

    def mc_volume(self, volume, X, Y, Z, grid=2):
        vol = np.ascontiguousarray(volume, np.float32)
        X, Y, Z = (np.ascontiguousarray(a, np.float64) for a in (X, Y, Z))
        assert vol.shape == (len(X), len(Y), len(Z))
        cap = max(1, (len(X) - 1) * (len(Y) - 1) * (len(Z) - 1)) * 12
        out = np.empty((cap, 9), np.float32)
        n_tri = ctypes.c_longlong(0)
        rc = self.lib.hs_mc_volume(vol.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(X)), ctypes.c_int(len(Y)),
                                   ctypes.c_int(len(Z)), X.ctypes.data_as(ctypes.c_void_p), Y.ctypes.data_as(ctypes.c_void_p),
                                   Z.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p),
                                   ctypes.c_longlong(cap), ctypes.byref(n_tri), ctypes.c_int(grid))
        assert rc == 0 and n_tri.value <= cap, (rc, n_tri.value, cap)
        return out[:n_tri.value]

    def stl_records(self, tris):
        tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
        out = np.zeros(len(tris) * 50, np.uint8)
        rc = self.lib.hs_stl_records(tris.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(len(tris)),
                                     out.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0, rc
        return out

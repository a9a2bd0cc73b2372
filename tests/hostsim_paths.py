"""Host-orchestration paths of sdfb200.cu that small lattices do not reach on their own, driven on the host-compiled
library (run by tests/test_hostsim_library_cpu.py in a child process started with --hostsim; not collected directly).

  * rounds cut by the HOST: tile volumes beyond the budget (24 GB on the GPU, a few tiles here)
  * the retry after a work-list overflow (lists start at 1 M entries on the GPU, at 64 here)
  * host output pipelined over device-cut rounds (SDFB200_PIPE): progress words, partial copies
  * device-resident output as ONE captured and replayed graph per parameter set (cache of four), count-only calls,
    Z-slabs of whole batch layers, the fused root gather played by three in-process ranks
  * the default specialisation mode: interpreter kernels while the compiler runs in the background
"""
import os

import numpy as np
import pytest

import models
from util_cmp import compare_soups

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def case():
    import sdf_b200 as S
    from oracle import sdf_ref
    from sdf_b200 import core
    f = models.cfg_csg(S)
    bounds, step = ((-1, -1, -1), (1, 1, 1)), 2.0 / 96
    _, _, (X, Y, Z) = core.make_lattice(f, step, bounds)
    h, _ = core.program_handle(f)
    tau = core.mixed_tau(h.program, X, Y, Z)
    rpts, rst = sdf_ref.generate(f, step=step, bounds=bounds)
    return dict(S=S, f=f, h=h, X=X, Y=Y, Z=Z, tau=tau, rpts=rpts, rst=rst, step=step, bounds=bounds)


def _run(case, precision='exact', zb=(0, -1)):
    h = case['h']
    n, stats, _ = h.generate(case['X'], case['Y'], case['Z'], 32, True, precision, case['tau'], None, 0, zb=zb)
    out = np.empty((n + 8, 9), np.float32)
    out[n:] = -777.0
    n2, stats2, _ = h.generate(case['X'], case['Y'], case['Z'], 32, True, precision, case['tau'], out, n, zb=zb)
    assert n2 == n and tuple(stats2[:3]) == tuple(stats[:3]) and np.all(out[n:] == -777.0)
    return out[:n], stats2


def _same_as_oracle(case, tris, stats):
    rst = case['rst']
    assert tuple(int(v) for v in stats[:3]) == (rst['skipped'], rst['empty'], rst['nonempty']), (stats, rst)
    rep = compare_soups(tris.reshape(-1, 3), case['rpts'])
    assert rep['count_equal'] and rep['exact_triangles'] == rep['n_ref'] and rep['n_ref'] > 50000, rep


def _debug_lines(capfd):
    """The library's SDFB200_DEBUG line per attempt: counters after the last round of the attempt."""
    import re
    out = []
    for line in capfd.readouterr().err.splitlines():
        m = re.match(r'\[sdfb200\] n_active=(\d+) n_slots=(\d+) n_cubes=(\d+) n_pts=(\d+) max=(\d+)/(\d+) caps=(\d+)/(\d+)', line)
        if m:
            out.append(dict(zip(('n_active', 'n_slots', 'n_cubes', 'n_pts', 'max_cubes', 'max_pts', 'cube_cap', 'pt_cap'),
                                (int(v) for v in m.groups()))))
    return out


@pytest.mark.parametrize('precision', ['exact', 'f64'])
def test_host_cut_rounds(case, precision, monkeypatch, capfd):
    """Tile volumes beyond the budget: the host reads the active count after the pre-pass and runs rounds of
    `budget` tiles over the same lists (27 batches, 2 tiles per round here)."""
    monkeypatch.setenv('HOSTSIM_VOL_BUDGET', str(2 * 35968 * 4))
    monkeypatch.setenv('SDFB200_DEBUG', '1')
    tris, stats = _run(case, precision)
    _same_as_oracle(case, tris, stats)
    dbg = _debug_lines(capfd)
    assert dbg and all(d['n_slots'] <= 2 and d['n_active'] > 10 for d in dbg), dbg       # the last round held <= 2 tiles


def test_work_list_overflow_is_retried(case, monkeypatch, capfd):
    """Lists that start far too small: the call notices the high-water marks, grows the lists and starts over from a
    clean slate (the pre-pass results stay valid); the result is the one of a call that never overflowed."""
    monkeypatch.setenv('HOSTSIM_MIN_CUBES', '64')
    monkeypatch.setenv('SDFB200_DEBUG', '1')
    from sdf_b200 import _abi
    h = _abi.ProgramHandle(case['h'].program)    # a fresh handle: its device context has no lists yet
    c = dict(case, h=h)
    tris, stats = _run(c)
    _same_as_oracle(case, tris, stats)
    dbg = _debug_lines(capfd)
    assert dbg[0]['max_cubes'] > dbg[0]['cube_cap'] == 64, dbg                           # first attempt: overflow
    assert dbg[1]['max_cubes'] <= dbg[1]['cube_cap'] and dbg[1]['max_pts'] <= dbg[1]['pt_cap'], dbg
    assert len(dbg) == 3, dbg                                                             # the second call fits at once


@pytest.mark.parametrize('rounds', [1, 3, 16])
def test_pipelined_host_output(case, rounds, monkeypatch):
    """Host output of a big enough lattice is cut into rounds ON THE DEVICE; each round's extraction reports through
    mapped progress words and its triangles are copied while the next round runs."""
    monkeypatch.setenv('SDFB200_PIPE', str(rounds))
    S = case['S']
    f = models.cfg_gearlike(S)
    L, n = 2.5, 128                              # 2^21 samples: the smallest lattice that pipelines
    pts, info = S.generate(f, step=2 * L / n, bounds=((-L, -L, -L), (L, L, L)), verbose=False, return_info=True)
    from oracle import sdf_ref
    rpts, rst = sdf_ref.generate(f, step=2 * L / n, bounds=((-L, -L, -L), (L, L, L)))
    assert (info['skipped'], info['empty'], info['nonempty']) == (rst['skipped'], rst['empty'], rst['nonempty'])
    rep = compare_soups(pts, rpts)
    assert rep['count_equal'] and rep['exact_triangles'] == rep['n_ref'] and rep['n_ref'] > 30000, rep


def test_z_slabs_add_up(case):
    """Whole batch layers [zb0, zb1): the slabs' soups and statistics add up to the whole call's."""
    whole, st = _run(case)
    parts, tot = [], np.zeros(3, np.int64)
    for zb in ((0, 1), (1, 2), (2, 3)):
        t, s = _run(case, zb=zb)
        parts.append(t)
        tot += np.asarray(s[:3])
    assert tuple(tot) == tuple(st[:3])
    rep = compare_soups(np.concatenate(parts).reshape(-1, 3), whole.reshape(-1, 3))
    assert rep['count_equal'] and rep['exact_triangles'] == rep['n_ref'], rep


def test_fused_root_gather_protocol(case, monkeypatch):
    """The multi-GPU C-ABI surface with three "ranks" played by one process (the stand-in runtime has HOSTSIM_DEVICES
    devices and an IPC handle is the pointer itself): every rank's extraction reserves its range in the ROOT's buffer
    through the shared cursor and signals `done`; the root waits for all signals, reads the count, consumes, releases;
    two buffers alternate by step parity and a rank's step s waits for `ready[s & 1] == s`.  Four steps: both buffers
    used twice.  (Between real GPUs the same calls run in tools/gpu_multi_check.py, one process per GPU.)"""
    import ctypes
    from sdf_b200 import _abi
    monkeypatch.setenv('HOSTSIM_DEVICES', '3')
    world, cap = 3, 100000
    root = _abi.SharedOutput(cap, 0)
    shared = [root] + [_abi.SharedOutput(cap, r, handle=root.handle) for r in (1, 2)]
    handles = [_abi.ProgramHandle(case['h'].program) for _ in range(world)]
    whole, st = _run(case)
    for step in range(4):
        which = step & 1
        local = {}
        for r in (2, 1, 0):                                   # the root last: its wait finds every signal in place
            n_local, stats, _ = handles[r].generate_shared(case['X'], case['Y'], case['Z'], 32, True, 'exact', case['tau'],
                                                           shared[r], which, zb=(r, r + 1), device=r)
            local[r] = (n_local, tuple(int(v) for v in stats[:3]))
        total = root.wait(which, world)
        assert total == sum(n for n, _ in local.values()) == len(whole), (step, total, local)
        assert tuple(np.sum([s for _, s in local.values()], axis=0)) == tuple(int(v) for v in st[:3])
        buf = (ctypes.c_float * (total * 9)).from_address(root.buffer_ptr(which))
        got = np.frombuffer(buf, np.float32).reshape(-1, 9).copy()
        rep = compare_soups(got.reshape(-1, 3), whole.reshape(-1, 3))
        assert rep['count_equal'] and rep['exact_triangles'] == rep['n_ref'], (step, rep)
        root.release(which)


class _DeviceBuffer:
    """"Device" memory under the stand-in runtime: the triangle buffer of a SharedOutput (cudaMalloc'ed by the library)."""

    def __init__(self, triangles):
        from sdf_b200 import _abi
        self.shared = _abi.SharedOutput(triangles, 0)
        self.ptr = self.shared.buffer_ptr(0)

    def data_ptr(self):
        return self.ptr

    def read(self, n):
        import ctypes
        return np.frombuffer((ctypes.c_float * (n * 9)).from_address(self.ptr), np.float32).reshape(-1, 9).copy()


def _graphs(capfd):
    import re
    return [int(m) for m in re.findall(r'graph=(\d+)', capfd.readouterr().err)]


def test_device_resident_step_is_one_replayed_graph(case, monkeypatch, capfd):
    """Device-resident output: the whole step -- counter reset, pre-pass, A, B, C, counter read-back -- is captured once
    per set of launch parameters and replayed (the stand-in runtime records the captured operations, kernel parameters
    copied at capture time as CUDA does, and replays them); a changed parameter captures a new graph, at most four are
    kept; with capture refused the plain launches give the same triangles."""
    from sdf_b200 import _abi
    monkeypatch.setenv('SDFB200_DEBUG', '1')
    h = _abi.ProgramHandle(case['h'].program)
    X, Y, Z, tau = case['X'], case['Y'], case['Z'], case['tau']
    n, stats, _ = h.generate(X, Y, Z, 32, True, 'exact', tau, None, 0)
    dev = _DeviceBuffer(n + 16)
    _graphs(capfd)
    for _ in range(3):
        dev.read(n)[:] = 0
        n2, stats2, _ = h.generate(X, Y, Z, 32, True, 'exact', tau, dev, n + 16)
        assert n2 == n
        _same_as_oracle(case, dev.read(n), stats2)
    assert _graphs(capfd) == [2, 2, 2]                      # the count-only call's graph + this one, replayed
    seen = []
    for k in range(6):                                      # six more parameter sets: the cache holds four graphs
        n3, stats3, _ = h.generate(X, Y, Z, 32, True, 'exact', tau * (1.0 + 0.01 * (k + 1)), dev, n + 16)
        assert n3 == n
        _same_as_oracle(case, dev.read(n), stats3)
        seen += _graphs(capfd)
    assert seen == [3, 4, 4, 4, 4, 4], seen
    monkeypatch.setenv('HOSTSIM_NO_CAPTURE', '1')
    h2 = _abi.ProgramHandle(case['h'].program)
    n4, stats4, _ = h2.generate(X, Y, Z, 32, True, 'exact', tau, dev, n + 16)
    assert n4 == n and _graphs(capfd) == [0]
    _same_as_oracle(case, dev.read(n), stats4)


def test_default_mode_runs_on_the_interpreter_until_the_compiler_is_done(monkeypatch):
    """SDFB200_JIT=1 (the default): a big enough call starts the compiles in the background and runs on the interpreter
    kernels; once the specialised kernels are there later calls use them -- same triangles either way."""
    import time
    import sdf_b200 as S
    from oracle import sdf_ref
    from sdf_b200 import core
    monkeypatch.setenv('SDFB200_JIT', '1')
    f = S.sphere(1.1) - S.box(1.3).k(0.07)                   # a tree no other test compiles: nothing in the kernel cache
    L, n = 1.25, 128                                         # 2^21 samples: "big enough to repay the compile"
    kw = dict(step=2 * L / n, bounds=((-L,) * 3, (L,) * 3))
    rpts, _ = sdf_ref.generate(f, **kw)
    first = S.generate(f, verbose=False, **kw)
    h, _ = core.program_handle(f)
    deadline = time.time() + 120
    share = None
    while time.time() < deadline and share is None:
        try:
            share = h.jit_plan()[0]                          # waits for (and loads) the specialised tile kernel
        except Exception:
            time.sleep(0.5)
    later = S.generate(f, verbose=False, **kw)
    for pts in (first, later):
        rep = compare_soups(pts, rpts)
        assert rep['count_equal'] and rep['exact_triangles'] == rep['n_ref'] and rep['n_ref'] > 20000, rep


@pytest.mark.parametrize('jit', ['0', '2'])
def test_streamed_float32_evaluator_through_the_abi(case, monkeypatch, jit):
    """sdfb200_eval_f32 / _fast with host buffers (staged through the library) and with "device" buffers, interpreter
    and run-time specialised stream kernels (on the GPU these entry points are driven with torch tensors)."""
    import ctypes
    from oracle import sdf_ref
    from sdf_b200 import _abi
    monkeypatch.setenv('SDFB200_JIT', jit)
    h = _abi.ProgramHandle(case['h'].program)
    n = 5000
    P = models.random_points(n, seed=4).astype(np.float32)
    r = sdf_ref.evaluate(case['f'], P.astype(np.float64)).reshape(-1)
    out = np.empty(n, np.float32)
    h.eval_f32(P, out, n)
    assert np.max(np.abs(out - r)) < 1e-5
    h.eval_f32_fast(P, out, n)
    assert np.max(np.abs(out - r)) < 1e-4
    dev_in, dev_out = _DeviceBuffer((n * 3 + 8) // 9 + 1), _DeviceBuffer(n // 9 + 1)
    ctypes.memmove(dev_in.ptr, P.ctypes.data, P.nbytes)
    h.eval_f32(dev_in, dev_out, n)
    got = np.frombuffer((ctypes.c_float * n).from_address(dev_out.ptr), np.float32)
    assert np.max(np.abs(got - r)) < 1e-5
    # page-locking a caller's buffer and the plain device -> host copy (the multi-GPU host gather's two calls)
    buf = np.zeros(n, np.float32)
    _abi.host_register(buf.ctypes.data, buf.nbytes)
    try:
        rc = _abi.lib().sdfb200_copy_to_host(buf.ctypes.data, dev_out.ptr, buf.nbytes, 0, None)
        assert rc == 0 and np.array_equal(buf, got)
    finally:
        _abi.host_unregister(buf.ctypes.data)

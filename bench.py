#!/usr/bin/env python
"""bench.py -- headline benchmark of the mesh-generation hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json metric "Mpoints/sec SDF-eval + Mtris/sec at 512^3"): configs[2], the
examples/gearlike.py tree on a 512^3 lattice (bounds +-2.5, step 5/512), sparse skip on, one
B200.  A "step" is one full pass of the hot path over the lattice: skip pre-pass, fused SDF
evaluation, marching cubes, world mapping.  For N > 1 (torchrun, one rank per GPU) the lattice
edge grows as 512 * N^(1/3) (rounded to a batch multiple) so every GPU keeps 512^3 samples
("weak"); ranks take Z-slabs (windows of the Z-major batch order) balanced by active batches and
re-planned from measured per-rank times.  `value` times the hot path with every rank's triangles
resident in its own HBM ("excluding the final all-gather, which is reported
separately"); the `gather` object reports the same series with the fused root gather (all
extraction kernels writing into rank 0 over NVLink).  `strong_1024` is north_star's strong-scaling
case (knurling, fixed 1024^3 lattice) measured the same two ways.

Output: ONE JSON line on rank 0 (see the task contract): value = Mpoints/s of lattice samples
covered (device-resident output), e2e = same through the public generate() with host output,
roofline for the fused kernel (12 algorithmic bytes per evaluated sample), cpu_baseline = the
CPU oracle port (numpy closures + C Lewiner MC) on the same lattice.

--impl reference times the reference's own CPU path: the UNMODIFIED package installed in
baseline/_ref (its numpy closure tree, _skip, _worker and ThreadPool, core.py:84-150) on all host
threads, with the one call it cannot make here -- skimage.measure.marching_cubes, scikit-image is
not installable -- served by oracle/mc_lewiner.c through a shim module ("reference eval + restated
MC"; the eval+skip-only time, which is purely the reference's code, is reported next to it).  For N > 1 the
lattice is N times larger and a step is a bounded sample of it: a seeded 1/N of its batches through the
reference's _worker + ThreadPool, the time extrapolated by N (rank 0 only; the other ranks exit).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np  # noqa: E402

BATCH = 32


def workload(n_gpus, grid=None, model='gearlike'):
    import models
    base, L = models.CONFIG_GRIDS[model]
    if grid is None:
        grid = int(round(base * n_gpus ** (1.0 / 3.0) / BATCH)) * BATCH
    return dict(model=model, n=grid, L=L, step=2.0 * L / grid, bounds=((-L, -L, -L), (L, L, L)))


def workload_name(w):
    """The same string in both arms (ours and --impl reference)."""
    return '%s tree (reference examples/%s.py), %d^3 lattice, bounds +-%g, sparse skip on, batch 32' % (
        w['model'], w['model'], w['n'], w['L'])


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fp:
            return float(json.load(fp)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class NvmlSampler:
    """SM clock + throttle reasons polled through NVML every millisecond during the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.ok = index, [], False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.ok = True
        except Exception:
            self.ok = False

    def start(self):
        if not self.ok:
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.time(), sm, reasons))
            except Exception:
                pass
            time.sleep(0.001)

    def stop(self, t0, t1):
        if not self.ok:
            return None
        self.stop_flag = True
        self.thread.join(1.0)
        nv = self.nv
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows
        if not rows:
            return None
        sm = sorted(r[1] for r in rows)
        bits = 0
        for r in rows:
            bits |= r[2]
        names = {'hw_slowdown': getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8) if hasattr(nv, 'nvmlClocksEventReasonHwSlowdown') else 0x8,
                 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20, 'sw_power_cap': 0x4}
        reasons = [k for k, m in names.items() if bits & m]
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        return {'sm_mhz': float(sm[len(sm) // 2]), 'sm_max_mhz': float(mx) if mx else None, 'reasons': reasons,
                'samples': len(rows), 'source': 'nvml, 1 ms polling during the timed region'}


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (fallback)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(',')]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows]
        sm = sorted(float(r[0]) for r in rows if r[0].replace('.', '').isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k].lower().startswith('active') for r in rows)]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(rows)}


def cpu_port(w, seconds_budget=20.0, workers=1):
    """The oracle port on a bounded, seeded sample of batches of the same lattice."""
    import models
    import sdf_b200
    from oracle import sdf_ref
    f = models.CONFIGS[w['model']](sdf_b200)
    nb = (w['n'] + BATCH - 1) // BATCH
    total = nb ** 3
    rng = np.random.default_rng(0)
    # calibrate on 64 batches, then size the sample to the budget
    probe = np.sort(rng.choice(total, min(64, total), replace=False))
    t = time.time()
    sdf_ref.generate(f, step=w['step'], bounds=w['bounds'], batch_subset=probe, workers=workers)
    per_batch = (time.time() - t) / len(probe)
    m = int(min(total, max(64, seconds_budget / max(per_batch, 1e-6))))
    sample = np.sort(rng.choice(total, m, replace=False)) if m < total else np.arange(total)
    t = time.time()
    pts, st = sdf_ref.generate(f, step=w['step'], bounds=w['bounds'], batch_subset=sample, workers=workers)
    dt = time.time() - t
    covered = m * BATCH ** 3 if m < total else w['n'] ** 3
    return dict(value=covered / dt / 1e6, unit='Mpoints/s', cores=workers, kind='port',
                sample='%d of %d batches (seeded), %d evaluated samples, %d triangles, %.1f s; numpy closure '
                       'restatement + C Lewiner MC (skimage unavailable)' % (m, total, st['evaluated'], st['triangles'], dt),
                mtris_per_s=st['triangles'] / dt / 1e6, seconds=dt)


def import_reference():
    """The UNMODIFIED reference package from baseline/_ref (pip-installed there with --no-deps).
    Its one missing dependency, scikit-image, is not installable here, so ``skimage.measure.marching_cubes`` is
    served by the C restatement (oracle/libmc_oracle.so) through a shim module -- every other line that runs is the
    reference's own: the numpy closure tree, _skip, _worker, the ThreadPool (core.py:84-150)."""
    import types
    ref_dir = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref_dir, 'sdf')):
        raise ImportError('baseline/_ref/sdf is missing (python -m pip install --no-index --no-build-isolation --no-deps '
                          '--target baseline/_ref <copy of /root/reference>)')
    from oracle import sdf_ref

    def marching_cubes(volume, level=0.0, **kw):
        soup = sdf_ref.marching_cubes_soup(volume, level)       # raises ValueError / RuntimeError like skimage
        faces = np.arange(len(soup), dtype=np.int32).reshape(-1, 3)
        return soup, faces, None, None

    stub = types.ModuleType('skimage')
    stub.measure = types.ModuleType('skimage.measure')
    stub.measure.marching_cubes = marching_cubes
    stub.__version__ = 'shim: oracle/mc_lewiner.c'
    sys.modules['skimage'] = stub
    sys.modules['skimage.measure'] = stub.measure
    # the repo root holds an alias package named `sdf` (drop-in for scripts): the reference must win here
    for k in [k for k in sys.modules if k == 'sdf' or k.startswith('sdf.')]:
        del sys.modules[k]
    sys.path.insert(0, ref_dir)
    import sdf as ref
    assert os.path.dirname(os.path.abspath(ref.__file__)).startswith(ref_dir), ref.__file__
    return ref


def reference_sample_pass(ref, f, w, cores, frac, seed=20260923):
    """A pass over a seeded 1/frac of the workload's batches with the reference's own per-batch pipeline.

    The lattice and the batch list are built as core.generate builds them (core.py:110-119); every sampled batch
    goes through ``ref.core._worker`` on ``ThreadPool(cores).imap`` and the results are consumed as core.py:131-141
    consumes them (the Python-level ``points.extend`` included).  Returns (callable -> triangles in the sample,
    number of sampled batches)."""
    import itertools
    from functools import partial
    from multiprocessing.pool import ThreadPool
    (x0, y0, z0), (x1, y1, z1) = w['bounds']
    d = w['step']
    X, Y, Z = np.arange(x0, x1, d), np.arange(y0, y1, d), np.arange(z0, z1, d)
    cut = lambda A: [A[i:i + BATCH + 1] for i in range(0, len(A), BATCH)]      # noqa: E731
    batches = list(itertools.product(cut(X), cut(Y), cut(Z)))
    rng = np.random.default_rng(seed)
    pick = np.sort(rng.choice(len(batches), max(1, len(batches) // frac), replace=False))
    sample = [batches[i] for i in pick]
    work = partial(ref.core._worker, f, step=(d, d, d), sparse=True)

    def one_pass():
        points = []
        with ThreadPool(cores) as pool:
            for result in pool.imap(work, sample):
                if result is not None and len(result):
                    points.extend(result)
        return len(points) // 3
    return one_pass, len(sample)


def run_reference(args):
    """The reference's own multiprocessing CPU path on this box's host cores: its core.generate with
    workers = all host threads, on the same workload as our arm."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    w = workload(args.gpus, args.grid, args.model)
    import models
    cores = os.cpu_count() or 1
    try:
        ref = import_reference()
        kind, how = 'reference', ('unmodified reference package (baseline/_ref): numpy closures, _skip, _worker, '
                                  'ThreadPool(%d); skimage.measure.marching_cubes served by the C restatement '
                                  '(scikit-image not installable): "reference eval + restated MC"' % cores)
    except ImportError as e:
        ref, kind, how = None, 'port', 'baseline/_ref unavailable (%s): oracle port, ThreadPool(%d)' % (e, cores)
    times, tris, eval_only = [], 0, None
    n_batches = ((w['n'] + BATCH - 1) // BATCH) ** 3
    sample_note = 'the whole lattice per step (%d batches)' % n_batches
    if ref is not None:
        f = models.CONFIGS[w['model']](ref)
        if args.gpus > 1:
            # N > 1: the lattice is N times the single-GPU one; a step is a BOUNDED sample of it -- a seeded 1/N of
            # its batches through the reference's own _worker (core.py:45-60) and ThreadPool, consumed like
            # core.generate consumes them (core.py:131-141) -- and the step time is extrapolated by N
            one_pass, n_sample = reference_sample_pass(ref, f, w, cores, args.gpus)
            scale = n_batches / n_sample
            sample_note = ('a seeded sample of %d of %d batches per step through the reference\'s _worker + ThreadPool, '
                           'time extrapolated x %.3f' % (n_sample, n_batches, scale))
        else:
            def one_pass():
                return len(ref.generate(f, step=w['step'], bounds=w['bounds'], workers=cores, verbose=False)) // 3
            scale = 1.0
        for it in range(args.warmup + args.steps):
            t = time.time()
            n = one_pass()
            dt = (time.time() - t) * scale
            if it >= args.warmup:
                times.append(dt)
                tris = int(round(n * scale))
        # the same pass with marching cubes stubbed out: the part that is PURELY the reference's code
        saved = sys.modules['skimage.measure'].marching_cubes

        def no_mc(volume, level=0.0, **kw):
            raise RuntimeError('marching cubes stubbed out')
        ref.core.measure.marching_cubes = no_mc
        t = time.time()
        one_pass()
        eval_only = (time.time() - t) * scale
        ref.core.measure.marching_cubes = saved
    else:
        import sdf_b200
        from oracle import sdf_ref
        f = models.CONFIGS[w['model']](sdf_b200)
        for it in range(args.warmup + args.steps):
            t = time.time()
            pts, st = sdf_ref.generate(f, step=w['step'], bounds=w['bounds'], workers=cores)
            dt = time.time() - t
            if it >= args.warmup:
                times.append(dt)
                tris = st['triangles']
    covered = w['n'] ** 3
    ms = 1e3 * sum(times) / len(times)
    val = covered / (ms / 1e3) / 1e6
    line = {'impl': 'reference', 'metric': 'Mpoints/s', 'value': val, 'unit': 'Mpoints/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': workload_name(w), 'points': w['n'] ** 3, 'precision': 'f64 (the reference arithmetic)'},
            'cpu_baseline': {'value': val, 'unit': 'Mpoints/s', 'cores': cores, 'kind': kind,
                             'sample': '%s; %s' % (sample_note, how),
                             'eval_and_skip_only_mpoints_per_s': (covered / eval_only / 1e6) if eval_only else None},
            'mtris_per_s': tris / (ms / 1e3) / 1e6, 'triangles': tris,
            'e2e': {'value': val, 'unit': 'Mpoints/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


class DeviceSeries:
    """K timed passes of the hot path with device-resident output on this rank's Z-slab (barrier + synchronize on
    both sides, CUDA events per step on the launching stream, L2 flushed between steps, max over ranks)."""

    def __init__(self, f, w, args, world, rank, local, flush, gather='none', plan=None):
        """``gather``: 'none' = every rank keeps its slab's triangles in its own HBM (what `value` times: SURVEY.md
        section 8d, "excluding the final all-gather, which is reported separately"); 'root' = the fused gather, every
        extraction kernel writes into rank 0's buffer over NVLink; 'allgather' = NCCL send/recv after the step."""
        import torch
        import torch.distributed as dist
        from sdf_b200 import core, multi
        self.torch, self.dist, self.world, self.rank, self.args, self.flush = torch, dist, world, rank, args, flush
        self.f, self.w, self.gather = f, w, gather
        _, _, (X, Y, Z) = core.make_lattice(f, w['step'], w['bounds'])
        self.axes = (X, Y, Z)
        self.n_points = len(X) * len(Y) * len(Z)
        self.stream = torch.cuda.current_stream().cuda_stream
        self.plan = (plan if plan is not None else multi.plan_windows(f, X, Y, Z, BATCH, world, device=local)) if world > 1 else None
        self.h, _ = core.program_handle(f)
        self.tau = core.mixed_tau(self.h.program, X, Y, Z)
        self.local = local
        self.root_gather = None
        self._size_buffers()
        self.gather_events = []
        if world > 1 and not args.no_rebalance and plan is None:
            self.rebalance()

    def _size_buffers(self):
        """Count-only pass of this rank's window -> output capacity (and the shared root buffer, once)."""
        torch, dist, args = self.torch, self.dist, self.args
        X, Y, Z = self.axes
        self.window = self.plan[self.rank] if self.plan is not None else None
        nb = [(len(a) + BATCH - 1) // BATCH for a in (X, Y, Z)]
        win = self.window if self.window is not None else (0, nb[0] * nb[1] * nb[2])
        need, _, _ = self.h.generate_window(X, Y, Z, BATCH, True, args.precision, self.tau, win, device=self.local, stream=self.stream)
        self.win = win
        self.cap = int(need * 1.02) + 1024
        self.out = torch.empty((self.cap, 9), dtype=torch.float32, device='cuda')
        if self.world > 1 and self.gather == 'root' and self.root_gather is None:
            from sdf_b200 import multi
            needs = torch.tensor([need], dtype=torch.int64, device='cuda')
            dist.all_reduce(needs)
            self.root_gather = multi.RootGather(int(needs.item() * 1.05) + 65536, self.local)

    def rebalance(self, rounds=4, tolerance=1.04):
        """Measured re-plan (at most ``rounds`` times): three untimed steps, per-rank device time of the step, windows
        re-cut with the batches of slower ranks weighing more (multi.reweigh) until max/min <= ``tolerance``."""
        torch, dist = self.torch, self.dist
        from sdf_b200 import multi
        X, Y, Z = self.axes
        weights = multi.batch_weights(self.f, X, Y, Z, BATCH, device=self.local)
        self.replans = []
        for _ in range(rounds):
            ms = []
            for _ in range(3):
                _, _, timings, _ = self.step()
                ms.append(float(timings[2]))
            mine = torch.tensor([sorted(ms)[1]], dtype=torch.float64, device='cuda')
            allt = [torch.zeros_like(mine) for _ in range(self.world)]
            dist.all_gather(allt, mine)
            seconds = [float(t.item()) * 1e-3 for t in allt]
            self.replans.append([round(t * 1e3, 4) for t in seconds])
            if max(seconds) <= tolerance * min(seconds):
                break
            weights = multi.reweigh(weights, self.plan, seconds)
            self.plan = multi.plan_windows(self.f, X, Y, Z, BATCH, self.world, device=self.local, weights=weights)
            self._size_buffers()

    def step(self):
        torch, args = self.torch, self.args
        X, Y, Z = self.axes
        if self.root_gather is not None:
            return self.root_gather.run(self.h, X, Y, Z, BATCH, True, args.precision, self.tau, None, stream=self.stream,
                                        window=self.win)
        n_tri, stats, timings = self.h.generate_window(X, Y, Z, BATCH, True, args.precision, self.tau, self.win,
                                                       tri_out=self.out, capacity=self.cap, device=self.local, stream=self.stream)
        gathered = None
        if self.world > 1 and self.gather == 'allgather':
            from sdf_b200 import multi
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            gathered = multi.allgather_triangles(self.out[:n_tri])
            g1.record()
            self.gather_events.append((g0, g1))
        return n_tri, stats, timings, gathered

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run(self, steps, warmup, sampler=None):
        torch, dist, world = self.torch, self.dist, self.world
        for _ in range(warmup):
            self.step()
        self.barrier()
        if sampler is not None:
            sampler.start()
        t0 = time.time()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        fused_ms, prepass_ms, kern_ms, stats_acc, n_tri = [], [], [], None, 0
        for k in range(steps):
            self.flush.zero_()                                  # L2 flush between timed iterations (not timed)
            ev[k][0].record()
            n_tri, stats, timings, _ = self.step()
            ev[k][1].record()
            fused_ms.append(timings[1])
            prepass_ms.append(timings[0])
            kern_ms.append((timings[4], timings[5], timings[6], timings[7]))
            stats_acc = stats
        self.barrier()
        t1 = time.time()
        step_ms = [a.elapsed_time(b) for a, b in ev]
        gather_ms = [a.elapsed_time(b) for a, b in self.gather_events[-steps:]] if self.gather_events else []
        mine = sum(step_ms) / steps
        t_max = torch.tensor([mine, sum(fused_ms) / steps, sum(k[0] for k in kern_ms) / steps], dtype=torch.float64, device='cuda')
        t_min = t_max.clone()
        agg = torch.tensor([float(n_tri), float(stats_acc[3]), float(stats_acc[4]), float(stats_acc[5]),
                            float(stats_acc[0]), float(stats_acc[1]), float(stats_acc[2])], dtype=torch.float64, device='cuda')
        kmax = torch.tensor([sum(k[i] for k in kern_ms) / steps for i in range(4)] + [sum(prepass_ms) / steps], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
            dist.all_reduce(t_min, op=dist.ReduceOp.MIN)
            dist.all_reduce(agg, op=dist.ReduceOp.SUM)
            dist.all_reduce(kmax, op=dist.ReduceOp.MAX)
        return dict(ms_per_step=float(t_max[0].item()), fused_ms=float(t_max[1].item()), tile_ms=float(t_max[2].item()),
                    per_rank_min=dict(step_ms=float(t_min[0].item()), pipeline_ms=float(t_min[1].item()), tile_ms=float(t_min[2].item())),
                    per_rank_max=dict(step_ms=float(t_max[0].item()), pipeline_ms=float(t_max[1].item()), tile_ms=float(t_max[2].item())),
                    agg=[float(v) for v in agg.tolist()], prepass_ms=sum(prepass_ms) / steps,
                    kern_max=dict(zip(('sign_eval', 'refine', 'extract', 'classify', 'prepass+setup'), [float(v) for v in kmax.tolist()])),
                    kern_ms=[sum(k[i] for k in kern_ms) / steps for i in range(4)],
                    gather_ms=(sum(gather_ms) / len(gather_ms)) if gather_ms else None, t0=t0, t1=t1)


COLD_SCRIPT = r"""
import json, os, sys, time
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, 'tests'))
import numpy as np
import models, sdf_b200
f0 = sdf_b200.sphere(1)
f0(np.zeros((8, 3)))                                  # creates the CUDA context (not ours to speed up)
f = models.CONFIGS[%(model)r](sdf_b200)
out = {}
for label in ('first', 'second'):
    t = time.perf_counter()
    pts = sdf_b200.generate(f, step=%(step)r, bounds=%(bounds)r, verbose=False, device=%(dev)d)
    out[label + '_ms'] = (time.perf_counter() - t) * 1e3
out['triangles'] = int(len(pts) // 3)
print(json.dumps(out))
"""


def cold_start(w, local):
    """First generate() of this tree in a FRESH process with an EMPTY kernel cache, default settings: the
    specialised kernels compile in the background while the call runs on the interpreter kernels."""
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        env = dict(os.environ, SDFB200_JIT_CACHE=tmp)
        env.pop('SDFB200_JIT', None)
        code = COLD_SCRIPT % dict(root=ROOT, model=w['model'], step=w['step'], bounds=w['bounds'], dev=local)
        try:
            r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=240)
            return json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as e:           # never fail the bench line over the side measurement
            return {'error': str(e)[:200]}


def run_ours(args):
    os.environ.setdefault('SDFB200_JIT', '2')       # timed arms: always the specialised kernels, compiled before the warm-up ends
    import torch
    import torch.distributed as dist
    import models
    import sdf_b200
    from sdf_b200 import multi

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.float32, device='cuda')   # 768 MB > 126 MB L2
    w = workload(world, args.grid, args.model)
    f = models.CONFIGS[w['model']](sdf_b200)
    series = DeviceSeries(f, w, args, world, rank, local, flush)
    X, Y, Z = series.axes
    n_points, h, plan = series.n_points, series.h, series.plan

    sampler = None
    if rank == 0:
        sampler = NvmlSampler(local)
        if not sampler.ok:
            sampler = ClockSampler(local)
    res = series.run(args.steps, args.warmup, sampler)
    ms_per_step = res['ms_per_step']
    clocks = sampler.stop(res['t0'], res['t1']) if rank == 0 else None

    def barrier():
        series.barrier()

    # end to end through the public API: host axes in, host triangle array out (rank-local slab)
    e2e_times, d2h = [], 0
    for it in range(2 + args.steps):
        barrier()
        t = time.perf_counter()
        if world > 1:
            pts, info = multi.generate_distributed(f, step=w['step'], bounds=w['bounds'], precision=args.precision,
                                                   plan=plan, gather=args.e2e_gather)
        else:
            pts, info = sdf_b200.generate(f, step=w['step'], bounds=w['bounds'], verbose=False,
                                          precision=args.precision, device=local, return_info=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        if it >= 2:
            e2e_times.append(dt)
            d2h = int(pts.size * 4)
    e2e_t = torch.tensor([sum(e2e_times) / len(e2e_times)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t.item())

    # the same series with the triangles gathered: fused root gather (every extraction kernel writes into rank 0's HBM over
    # NVLink) -- reported separately, `value` is the slab-resident series
    gather = None
    if world > 1:
        sg = DeviceSeries(f, w, args, world, rank, local, flush, gather=args.gather, plan=series.plan)
        rg = sg.run(max(3, min(args.steps, 10)), 3)
        gather = {'kind': 'fused root gather: peer writes into rank 0 over NVLink, ready/done flag words, no collective' if args.gather == 'root'
                  else 'NCCL send/recv all-gather after the step',
                  'ms_per_step_with_gather': rg['ms_per_step'], 'value_with_gather': n_points / (rg['ms_per_step'] / 1e3) / 1e6,
                  'output_mb': rg['agg'][0] * 36 / 1e6, 'kernel_ms_max_over_ranks': rg['kern_max'],
                  'note': 'the whole soup lands in ONE GPU: bounded below by that GPU\'s NVLink ingress (output_mb / ~0.8 TB/s)'}
        del sg

    # north_star's STRONG-scaling case next to the weak series the --gpus contract asks for: the knurling tree on a
    # FIXED 1024^3 lattice split over the ranks (compare this object's ms_per_step across the N = 1, 2, 4, 8 lines)
    strong = None
    if not args.no_strong:
        ws = workload(1, 1024, 'knurling')
        fs = models.CONFIGS['knurling'](sdf_b200)
        ss = DeviceSeries(fs, ws, args, world, rank, local, flush)
        rs = ss.run(max(3, min(args.steps, 10)), 3)
        strong = {'workload': workload_name(ws), 'scaling': 'strong', 'n_gpus': world, 'ms_per_step': rs['ms_per_step'],
                  'value': ss.n_points / (rs['ms_per_step'] / 1e3) / 1e6, 'unit': 'Mpoints/s',
                  'mtris_per_s': rs['agg'][0] / (rs['ms_per_step'] / 1e3) / 1e6, 'triangles': int(rs['agg'][0]),
                  'per_rank_min_ms': rs['per_rank_min'], 'per_rank_max_ms': rs['per_rank_max'],
                  'windows': [list(p) for p in ss.plan] if ss.plan else None,
                  'replans_per_rank_ms': getattr(ss, 'replans', None), 'kernel_ms_max_over_ranks': rs['kern_max'],
                  'output': 'every rank keeps its slab in its own HBM (gather reported separately)'}
        if world > 1:
            sgs = DeviceSeries(fs, ws, args, world, rank, local, flush, gather=args.gather, plan=ss.plan)
            rgs = sgs.run(5, 3)
            strong['with_root_gather'] = {'ms_per_step': rgs['ms_per_step'], 'output_mb': rgs['agg'][0] * 36 / 1e6,
                                          'kernel_ms_max_over_ranks': rgs['kern_max']}
            del sgs
        del ss

    if rank == 0:
        tri_total, evaluated, refined, cubes = res['agg'][:4]
        peak, peak_src = peaks()
        tile_s = res['tile_ms'] / 1e3
        tile_kernel_name = ('tile_sign_body (jit_kernel, 592 x 256: fused SDF evaluation + sign masks)' if args.precision == 'exact' else 'tile_eval_body (jit_kernel, 148 x 1024)')
        achieved = (evaluated / world) * 12.0 / tile_s / 1e9 if tile_s > 0 else 0.0     # per-GPU kernel
        # DRAM traffic of the dominant kernel: an ncu figure (profiles/r2_traffic.json, written from the committed
        # capture of THIS kernel on THIS workload); reported only when that capture's kernel time agrees with the
        # time measured in this run -- a stale capture (another kernel version) prints null instead of a number
        traffic, traffic_note = None, 'no ncu capture for this configuration'
        tpath = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
        if os.path.exists(tpath) and world == 1 and args.grid is None:
            with open(tpath) as fp:
                rec = json.load(fp).get(args.model + ':' + args.precision)
            if rec:
                t_ncu, t_now = float(rec['kernel_us']), tile_s * 1e6
                if abs(t_ncu - t_now) <= 0.35 * t_now:
                    traffic = int(rec['dram_bytes_per_launch'])
                    traffic_note = 'ncu %s (%s): %.1f us per launch under ncu, %.1f us here' % (rec['capture'], rec['kernel'], t_ncu, t_now)
                else:
                    traffic_note = 'capture %s is stale: %.1f us under ncu vs %.1f us measured here' % (rec['capture'], t_ncu, t_now)
        cold = cold_start(w, local) if (world == 1 and not args.no_cold) else None
        line = {
            'metric': 'Mpoints/s', 'value': n_points / (ms_per_step / 1e3) / 1e6, 'unit': 'Mpoints/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': {'mixed': 'f32 (+f64 re-evaluation of |d|<tau samples)', 'f32': 'f32', 'f64': 'f64', 'exact': 'f32 classification + f64 for every sample that reaches the output'}[args.precision],
            'data': 'synthetic',
            'config': {'workload': workload_name(w),
                       'points': n_points, 'precision': args.precision,
                       'l2': 'flushed between timed iterations (768 MB memset, outside the per-step events)',
                       'parallelism': 'z-slab x%d (windows of the Z-major batch order, balanced by active batches%s)' % (
                           world, ', measured re-plans' if (world > 1 and not args.no_rebalance) else ''),
                       'gather': ('excluded from value (every rank keeps its slab; SURVEY.md 8d) -- see the `gather` object' if world > 1 else None),
                       'sample_grouping': dict(zip(('share_axis', 'samples_per_thread', 'invariant_fraction'),
                                                   h.jit_plan({'f32': 0, 'mixed': 1, 'f64': 2, 'exact': 3}[args.precision]))),
                       'host_output_rounds': int(os.environ.get('SDFB200_PIPE', 8))},
            'mtris_per_s': tri_total / (ms_per_step / 1e3) / 1e6, 'triangles': int(tri_total),
            'mpoints_evaluated_per_s': evaluated / (ms_per_step / 1e3) / 1e6,
            'sdf_eval_mpoints_per_s': (evaluated / world) / tile_s / 1e6 if tile_s > 0 else None,
            'batches': {'skipped': int(res['agg'][4]), 'empty': int(res['agg'][5]), 'nonempty': int(res['agg'][6])},
            'samples_evaluated': int(evaluated), 'samples_refined_f64': int(refined), 'active_cubes': int(cubes),
            'kernel_ms': {'tile_pipeline_A+B+C': res['fused_ms'], 'skip_prepass+setup': res['prepass_ms'],
                          'tile_kernel(A)': res['kern_ms'][0], 'tile_classify_kernel(A2)': res['kern_ms'][3],
                          'refine_points_kernel(B)': res['kern_ms'][1],
                          'extract_kernel(C)': res['kern_ms'][2],
                          'per_rank_min': res['per_rank_min'], 'per_rank_max': res['per_rank_max']},
            'gpu_launches': 5 * args.steps,
            'clocks': clocks,
            'e2e': {'value': n_points / e2e_s / 1e6, 'unit': 'Mpoints/s', 'h2d_bytes_per_step': int(8 * (len(X) + len(Y) + len(Z))),
                    'd2h_bytes_per_step': d2h, 'ms_per_step': e2e_s * 1e3, 'mtris_per_s': tri_total / e2e_s / 1e6,
                    'd2h_gbs_if_copy_bound': d2h / e2e_s / 1e9,
                    'cold_ms': (cold or {}).get('first_ms'), 'cold': cold},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'traffic_source': traffic_note,
                         'kernel': tile_kernel_name, 'peak_source': peak_src,
                         'algorithmic_bytes_per_launch': (evaluated / world) * 12.0,
                         'note': 'dominant kernel of the step (tile evaluate + classify); 12 algorithmic B per evaluated '
                                 'sample (north_star float3 accounting) over its CUDA-event duration, measured live in this '
                                 'run; the kernel generates its samples in registers from the three lattice axes, so its real '
                                 'DRAM traffic is far below the algorithmic bytes and it is FP32/MUFU issue bound'},
            'strong_1024': strong,
            'gather': gather,
        }
        if not args.no_cpu_baseline and world == 1:
            line['cpu_baseline'] = cpu_port(w, args.cpu_seconds)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--model', default='gearlike')
    ap.add_argument('--grid', type=int, default=None)
    ap.add_argument('--precision', default='exact', choices=['f32', 'mixed', 'f64', 'exact'])
    ap.add_argument('--cpu-seconds', type=float, default=15.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-strong', action='store_true', help='skip the knurling 1024^3 strong-scaling side series')
    ap.add_argument('--no-cold', action='store_true', help='skip the cold-start side measurement')
    ap.add_argument('--e2e-gather', default='shm', choices=['shm', 'root', 'host'],
                    help='N>1, public API leg: every rank copies its slab to one shared pinned host segment over its own PCIe link (default), '
                         'fused root gather + one copy, or NCCL all-gather + copy')
    ap.add_argument('--no-rebalance', action='store_true', help='N>1: keep the plan balanced by active batches (no measured re-plan)')
    ap.add_argument('--gather', default='root', choices=['root', 'allgather'],
                    help='N>1: fused NVLink peer writes into rank 0 (default) or NCCL send/recv all-gather')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()

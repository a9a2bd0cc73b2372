"""A/B timing on a B200 of the two speed-only mechanisms of the tile pipeline (run under gpurun):

  * sample grouping of the specialised tile kernel: flat 4-sample mapping vs the automatic choice
    (axis sharing, sdfb200_jit_plan), kernel times from the C-ABI's CUDA-event timings, device output;
  * pipelined host output (SDFB200_PIPE rounds): wall time of the public generate() with host output.

Every variant also reports the triangle count and an order-independent checksum of the soup, so the
lines prove that the variants produce the same mesh.  One JSON object per line on stdout.

  python tools/gpu_ab.py [--models gearlike,knurling,...] [--quick]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np  # noqa: E402


def soup_checksum(tri9):
    """Order-independent 64-bit checksum of a (T, 9) float32 soup (rows hashed, hashes summed)."""
    u = np.ascontiguousarray(tri9).view(np.uint32).reshape(-1, 9).astype(np.uint64)
    h = np.full(len(u), 0xcbf29ce484222325, np.uint64)
    with np.errstate(over='ignore'):
        for k in range(9):
            h = (h ^ u[:, k]) * np.uint64(0x100000001b3)
            h ^= h >> np.uint64(29)
        return int(h.sum(dtype=np.uint64))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--models', default='gearlike,knurling,csg,blobby,sphere')
    ap.add_argument('--quick', action='store_true')
    ap.add_argument('--grid', type=int, default=0)
    ap.add_argument('--variants', default='flat4,auto',
                    help='comma list of: flat4, auto, col11 (z columns of 11 whatever the plan says), gather '
                         '(SDF_EXP_REFM_GATHER), minmax (SDF_EXP_NUMPY_MINMAX), carveout (SDFB200_EXP_CARVEOUT), '
                         'and combinations joined with + (e.g. auto+gather+carveout)')
    ap.add_argument('--no-host-output', action='store_true')
    args = ap.parse_args()
    os.environ['SDFB200_JIT'] = '2'
    import torch
    import models
    import sdf_b200
    from sdf_b200 import _abi, core
    from sdf_b200.ir import lower

    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device='cuda')
    reps = 4 if args.quick else 10
    for name in args.models.split(','):
        n, L = models.CONFIG_GRIDS[name]
        if args.grid:
            n = args.grid
        f = models.CONFIGS[name](sdf_b200)
        bounds, step = ((-L, -L, -L), (L, L, L)), 2.0 * L / n
        _, _, (X, Y, Z) = core.make_lattice(f, step, bounds)
        prog = lower(f)
        tau = core.mixed_tau(prog, X, Y, Z)
        out = None
        for label in args.variants.split(','):
            parts = label.split('+')
            share, group = ('-1', '4') if 'flat4' in parts else (('2', '11') if 'col11' in parts else (None, None))
            defines = ' '.join(d for d, key in (('SDF_EXP_REFM_GATHER', 'gather'), ('SDF_EXP_NUMPY_MINMAX', 'minmax')) if key in parts)
            for k, v in (('SDFB200_JIT_SHARE', share), ('SDFB200_JIT_N', group), ('SDFB200_JIT_DEFINES', defines or None),
                         ('SDFB200_EXP_CARVEOUT', '1' if 'carveout' in parts else None)):
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
            h = _abi.ProgramHandle(prog)
            plan = h.jit_plan(3)
            t0 = time.time()
            need, _, _ = h.generate(X, Y, Z, 32, True, 'exact', tau, None, 0, device=0, stream=stream)
            jit_s = time.time() - t0
            if out is None:
                cap = int(need * 1.02) + 1024
                out = torch.empty((cap, 9), dtype=torch.float32, device='cuda')
            acc = []
            for it in range(2 + reps):
                flush.zero_()
                n_tri, stats, tm = h.generate(X, Y, Z, 32, True, 'exact', tau, out, cap, device=0, stream=stream)
                if it >= 2:
                    acc.append([tm[0], tm[1], tm[4], tm[5], tm[6]])
            a = np.array(acc)
            med = np.median(a, axis=0)
            print(json.dumps(dict(kind='grouping', model=name, grid=n, variant=label, share_axis=plan[0], group=plan[1],
                                  invariant_fraction=round(plan[2], 3), first_call_s=round(jit_s, 2), triangles=int(n_tri),
                                  checksum=soup_checksum(out[:n_tri].cpu().numpy()), evaluated=int(stats[3]), refined=int(stats[4]),
                                  prepass_ms=round(float(med[0]), 4), pipeline_ms=round(float(med[1]), 4),
                                  tile_eval_ms=round(float(med[2]), 4), refine_ms=round(float(med[3]), 4),
                                  extract_ms=round(float(med[4]), 4), tile_eval_min_ms=round(float(a[:, 2].min()), 4))), flush=True)
            del h
        # pipelined host output through the public API (automatic grouping)
        for k in ('SDFB200_JIT_SHARE', 'SDFB200_JIT_N', 'SDFB200_JIT_DEFINES', 'SDFB200_EXP_CARVEOUT'):
            os.environ.pop(k, None)
        for rounds in (() if args.no_host_output else ('1', '2', '4', '8')):
            os.environ['SDFB200_PIPE'] = rounds
            ts = []
            for it in range(2 + reps):
                torch.cuda.synchronize()
                t = time.perf_counter()
                pts, info = sdf_b200.generate(f, step=step, bounds=bounds, verbose=False, precision='exact', device=0,
                                              return_info=True)
                torch.cuda.synchronize()
                if it >= 2:
                    ts.append(time.perf_counter() - t)
            print(json.dumps(dict(kind='host_output', model=name, grid=n, pipe_rounds=int(rounds), triangles=len(pts) // 3,
                                  checksum=soup_checksum(pts.reshape(-1, 9)), e2e_ms_median=round(float(np.median(ts)) * 1e3, 4),
                                  e2e_ms_min=round(min(ts) * 1e3, 4), mpoints_per_s=round(len(X) * len(Y) * len(Z) / float(np.median(ts)) / 1e6, 1))),
                  flush=True)
        os.environ.pop('SDFB200_PIPE', None)
        del out


if __name__ == '__main__':
    main()

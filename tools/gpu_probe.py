#!/usr/bin/env python
"""Timing probe: every BASELINE config at full size, per precision mode (GPU box only)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import models
import sdf_b200
from sdf_b200 import core


def main():
    names = sys.argv[1].split(',') if len(sys.argv) > 1 else list(models.CONFIGS)
    precs = sys.argv[2].split(',') if len(sys.argv) > 2 else ['f32', 'mixed', 'exact', 'f64']
    for name in names:
        n, L = models.CONFIG_GRIDS[name]
        f = models.CONFIGS[name](sdf_b200)
        _, _, (X, Y, Z) = core.make_lattice(f, 2.0 * L / n, ((-L,) * 3, (L,) * 3))
        h, _ = core.program_handle(f)
        tau = core.mixed_tau(h.program, X, Y, Z)
        for prec in precs:
            need, _, _ = h.generate(X, Y, Z, 32, True, prec, tau, None, 0)
            out = torch.empty((need + 16, 9), dtype=torch.float32, device='cuda')
            best = None
            for it in range(6):
                n_tri, stats, t = h.generate(X, Y, Z, 32, True, prec, tau, out, need + 16)
                if best is None or t[2] < best[2]:
                    best = t.copy()
            print(json.dumps(dict(model=name, n=n, prec=prec, code_words=len(h.program.code), tri=int(n_tri),
                                  skipped=int(stats[0]), empty=int(stats[1]), nonempty=int(stats[2]),
                                  evaluated=int(stats[3]), refined=int(stats[4]), cubes=int(stats[5]),
                                  prepass_ms=round(best[0], 4), fused_ms=round(best[1], 4), tile_ms=round(best[4], 4),
                                  classify_ms=round(best[7], 4), refine_ms=round(best[5], 4), extract_ms=round(best[6], 4), plan=h.jit_plan(core._abi.PREC[prec]),
                                  frac_12B=round(stats[3] * 12.0 / (best[4] * 1e-3) / 6570.3e9, 4) if best[4] > 0 else None,
                                  gpts_eval_per_s=round(stats[3] / best[1] / 1e6, 2),
                                  gpts_covered_per_s=round(n ** 3 / best[2] / 1e6, 2),
                                  mtris_per_s=round(n_tri / best[2] / 1e3, 1))), flush=True)
            del out
    if os.environ.get('PROBE_NO_STREAM'):
        return
    # streamed evaluator: 512^3 float3 points from HBM
    npts = 512 ** 3
    rng = torch.Generator(device='cuda').manual_seed(0)
    P = (torch.rand((npts, 3), device='cuda', generator=rng) * 5 - 2.5)
    for name, fast in [(n, fa) for n in names for fa in (False, True)]:
        f = models.CONFIGS[name](sdf_b200)
        h, _ = core.program_handle(f)
        out = torch.empty(npts, device='cuda')
        s = torch.cuda.current_stream().cuda_stream
        fn = h.eval_f32_fast if fast else h.eval_f32
        for _ in range(2):
            fn(P, out, npts, stream=s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn(P, out, npts, stream=s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(json.dumps(dict(stream_eval=name, fast=fast, ms=round(ms, 3), gpts_per_s=round(npts / ms / 1e6, 2),
                              gbs_12B=round(npts * 12 / ms / 1e6, 1), gbs_16B=round(npts * 16 / ms / 1e6, 1))), flush=True)


if __name__ == '__main__':
    main()

#!/bin/bash
# Round 2, fifth GPU call (1 GPU): whole GPU suite (closure route, weld, guard band included), bench, reference arm.
mkdir -p gpurun_out
echo "== smoke"
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2e_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2e_smoke.log
echo "== pytest gpu (all)"
timeout 1500 python -m pytest tests -m "gpu" -q --timeout 600 > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r2e_pytest_gpu.log
echo "== bench"
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2e_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2e_bench.json').read().strip().splitlines()[-1])
    print('value %.0f ms %.4f e2e %.0f (%.3f ms) cold %s kernels %s roofline %.0f %.3f traffic %s strong %s' % (
        d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('cold'), d['kernel_ms'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['traffic'],
        {k: d['strong_1024'][k] for k in ('ms_per_step', 'value', 'triangles')} if d.get('strong_1024') else None))
except Exception as e:
    print('bench parse failed', e)
PY
echo "== e2e default pipe"
for m in gearlike blobby knurling csg; do timeout 100 python gpurun_out/_e2e.py $m; done | tee gpurun_out/r2e_e2e_default.jsonl
echo "== weld timing"
python - <<'PY'
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, models, sdf_b200
from sdf_b200 import core
for name in ('gearlike', 'knurling'):
    n, L = models.CONFIG_GRIDS[name]
    f = models.CONFIGS[name](sdf_b200)
    pts = sdf_b200.generate(f, step=2.0 * L / n, bounds=((-L,) * 3, (L,) * 3), verbose=False)
    core.weld(pts[:3000])
    t = time.perf_counter(); v, fc = core.weld(pts); dt = time.perf_counter() - t
    t = time.perf_counter(); rv, ri = np.unique(np.asarray(pts), axis=0, return_inverse=True); dn = time.perf_counter() - t
    print(name, 'triangles', len(pts) // 3, 'verts', len(v), 'gpu weld %.1f ms' % (dt * 1e3), 'np.unique %.1f ms' % (dn * 1e3), 'equal', np.array_equal(v, rv) and np.array_equal(fc.reshape(-1), ri.reshape(-1)))
PY

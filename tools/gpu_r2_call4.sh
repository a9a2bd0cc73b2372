#!/bin/bash
# Round 2, fourth GPU call (1 GPU): device-cut pipelined rounds for host output (progress words polled by the host),
# smaller sign batches with a prefetched cursor.
mkdir -p gpurun_out
echo "== smoke"
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2d_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2d_smoke.log
echo "== pytest gpu (not fullsize)"
timeout 900 python -m pytest tests -m "gpu and not fullsize" -x -q --timeout 300 > gpurun_out/r2d_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2d_pytest_gpu.log
echo "== pytest fullsize"
timeout 900 python -m pytest tests -m "fullsize" -q --timeout 600 > gpurun_out/r2d_pytest_fullsize.log 2>&1; echo "fullsize rc=$?"
tail -4 gpurun_out/r2d_pytest_fullsize.log
echo "== probe"
PROBE_NO_STREAM=1 timeout 300 python tools/gpu_probe.py sphere,csg,gearlike,blobby,knurling exact > gpurun_out/r2d_probe.jsonl 2> gpurun_out/r2d_probe.err; echo "probe rc=$?"
cat gpurun_out/r2d_probe.jsonl | cut -c1-640
echo "== e2e variants"
cat > gpurun_out/_e2e.py <<'PY'
import os, sys, time, json
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch, models, sdf_b200
os.environ['SDFB200_JIT'] = '2'
name = sys.argv[1]; n, L = models.CONFIG_GRIDS[name]
f = models.CONFIGS[name](sdf_b200)
ts = []
ck = None
for it in range(14):
    t = time.perf_counter()
    pts = sdf_b200.generate(f, step=2.0 * L / n, bounds=((-L,) * 3, (L,) * 3), verbose=False)
    ts.append(time.perf_counter() - t)
    if it == 3:
        ck = (len(pts) // 3, float(np.asarray(pts, np.float64).sum()))
ts = sorted(ts[3:])
print(json.dumps(dict(model=name, pipe=os.environ.get('SDFB200_PIPE'), graph=os.environ.get('SDFB200_GRAPH'), tri=ck[0], checksum=ck[1],
                      e2e_ms_median=round(ts[len(ts) // 2] * 1e3, 4), e2e_ms_min=round(ts[0] * 1e3, 4), d2h_mb=round(pts.nbytes / 1e6, 1))))
PY
for p in 1 2 3 4 6 8; do SDFB200_PIPE=$p timeout 100 python gpurun_out/_e2e.py gearlike; done | tee gpurun_out/r2d_e2e_variants.jsonl
SDFB200_PIPE=4 SDFB200_GRAPH=0 timeout 100 python gpurun_out/_e2e.py gearlike | tee -a gpurun_out/r2d_e2e_variants.jsonl
for p in 1 4 8; do SDFB200_PIPE=$p timeout 100 python gpurun_out/_e2e.py knurling; done | tee -a gpurun_out/r2d_e2e_variants.jsonl
for p in 1 4 8; do SDFB200_PIPE=$p timeout 100 python gpurun_out/_e2e.py blobby; done | tee -a gpurun_out/r2d_e2e_variants.jsonl
for p in 1 4; do SDFB200_PIPE=$p timeout 100 python gpurun_out/_e2e.py csg; done | tee -a gpurun_out/r2d_e2e_variants.jsonl
echo "== bench"
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2d_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2d_bench.json').read().strip().splitlines()[-1])
    print('value %.0f ms %.4f e2e %.0f (%.3f ms) cold %s kernels %s roofline %.0f %.3f strong %s' % (
        d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('cold'), d['kernel_ms'], d['roofline']['achieved'], d['roofline']['frac'],
        {k: d['strong_1024'][k] for k in ('ms_per_step', 'value', 'triangles')} if d.get('strong_1024') else None))
except Exception as e:
    print('bench parse failed', e)
PY
echo "== ncu full (tile, gearlike + blobby)"
cat > gpurun_out/_one.py <<'PY'
import os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch, models, sdf_b200
from sdf_b200 import _abi, core
from sdf_b200.ir import lower
name, n = sys.argv[1], int(sys.argv[2])
_, L = models.CONFIG_GRIDS[name]
f = models.CONFIGS[name](sdf_b200)
_, _, (X, Y, Z) = core.make_lattice(f, 2.0 * L / n, ((-L, -L, -L), (L, L, L)))
prog = lower(f); h = _abi.ProgramHandle(prog); tau = core.mixed_tau(prog, X, Y, Z)
s = torch.cuda.current_stream().cuda_stream
need, _, _ = h.generate(X, Y, Z, 32, True, 'exact', tau, None, 0, device=0, stream=s)
out = torch.empty((int(need * 1.02) + 1024, 9), dtype=torch.float32, device='cuda')
for _ in range(3):
    print(h.generate(X, Y, Z, 32, True, 'exact', tau, out, out.shape[0], device=0, stream=s)[0], h.jit_plan(3))
PY
cap() {   # name model grid kernel-regex launches-to-skip
  SDFB200_JIT=2 SDFB200_GRAPH=0 timeout 200 ncu --set full --clock-control none --import-source on \
    -k regex:$4 -s $5 -c 1 -f -o gpurun_out/$1 python gpurun_out/_one.py $2 $3 > gpurun_out/$1.log 2>&1
  echo "$1 rc=$? $(tail -1 gpurun_out/$1.log)"
}
cap r2d_tile_gearlike  gearlike 512 jit_kernel 10
cap r2d_tile_blobby    blobby 768 jit_kernel 10
ls -la gpurun_out/r2d*.ncu-rep

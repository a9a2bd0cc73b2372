#!/bin/bash
# Round 2, second GPU call: autonomous-warp sign kernel, one-graph step, bulk-copied MC tables, fused statistics.
mkdir -p gpurun_out
echo "== smoke"
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2b_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2b_smoke.log
echo "== pytest gpu (not fullsize)"
timeout 900 python -m pytest tests -m "gpu and not fullsize" -x -q --timeout 300 > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2b_pytest_gpu.log
echo "== pytest fullsize"
timeout 900 python -m pytest tests -m "fullsize" -q --timeout 600 > gpurun_out/r2b_pytest_fullsize.log 2>&1; echo "fullsize rc=$?"
tail -4 gpurun_out/r2b_pytest_fullsize.log
echo "== sanitizer (smoke: memcheck, racecheck)"
SDFB200_JIT=2 timeout 300 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/r2b_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|smoke ok" gpurun_out/r2b_memcheck.log
SDFB200_JIT=2 timeout 400 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/r2b_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|smoke ok" gpurun_out/r2b_racecheck.log
echo "== probe"
PROBE_NO_STREAM=1 timeout 300 python tools/gpu_probe.py sphere,csg,gearlike,blobby,knurling exact > gpurun_out/r2b_probe.jsonl 2> gpurun_out/r2b_probe.err; echo "probe rc=$?"
cat gpurun_out/r2b_probe.jsonl | cut -c1-620
echo "== probe no graph"
SDFB200_GRAPH=0 PROBE_NO_STREAM=1 timeout 300 python tools/gpu_probe.py gearlike exact > gpurun_out/r2b_probe_nograph.jsonl 2>&1; cat gpurun_out/r2b_probe_nograph.jsonl | cut -c1-620
echo "== bench"
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2b_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
    print('value %.0f ms %.4f e2e %.0f (%.3f ms) cold %s kernels %s roofline %.0f %.3f strong %s' % (
        d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('cold'), d['kernel_ms'], d['roofline']['achieved'], d['roofline']['frac'],
        {k: d['strong_1024'][k] for k in ('ms_per_step', 'value', 'triangles')} if d.get('strong_1024') else None))
except Exception as e:
    print('bench parse failed', e)
PY
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err; echo "ref rc=$?"; cut -c1-900 gpurun_out/r2b_bench_ref.json
echo "== pcie + e2e variants"
cat > gpurun_out/_e2e.py <<'PY'
import os, sys, time, json
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch, models, sdf_b200
os.environ['SDFB200_JIT'] = '2'
if len(sys.argv) > 1 and sys.argv[1] == 'pcie':
    for mb in (4, 32, 256):
        n = mb * 1024 * 1024 // 4
        d = torch.empty(n, dtype=torch.float32, device='cuda'); hbuf = torch.empty(n, dtype=torch.float32, pin_memory=True)
        for _ in range(3): hbuf.copy_(d, non_blocking=True)
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(10): hbuf.copy_(d, non_blocking=True)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
        print(json.dumps(dict(pcie_d2h_mb=mb, gbs=round(mb / 1024 / dt, 2), us=round(dt * 1e6, 1))))
    sys.exit(0)
name = sys.argv[1]; n, L = models.CONFIG_GRIDS[name]
f = models.CONFIGS[name](sdf_b200)
ts = []
for it in range(12):
    t = time.perf_counter()
    pts = sdf_b200.generate(f, step=2.0 * L / n, bounds=((-L,) * 3, (L,) * 3), verbose=False)
    ts.append(time.perf_counter() - t)
ts = sorted(ts[2:])
print(json.dumps(dict(model=name, pipe=os.environ.get('SDFB200_PIPE'), zerocopy=os.environ.get('SDFB200_ZEROCOPY'), tri=len(pts) // 3,
                      e2e_ms_median=round(ts[len(ts) // 2] * 1e3, 4), e2e_ms_min=round(ts[0] * 1e3, 4), d2h_mb=round(pts.nbytes / 1e6, 1))))
PY
timeout 60 python gpurun_out/_e2e.py pcie | tee gpurun_out/r2b_pcie.jsonl
for v in "8 0" "4 0" "16 0" "2 0" "1 0" "1 1"; do set -- $v; SDFB200_PIPE=$1 SDFB200_ZEROCOPY=$2 timeout 100 python gpurun_out/_e2e.py gearlike; done | tee gpurun_out/r2b_e2e_variants.jsonl
SDFB200_PIPE=8 timeout 100 python gpurun_out/_e2e.py knurling | tee -a gpurun_out/r2b_e2e_variants.jsonl
SDFB200_PIPE=16 timeout 100 python gpurun_out/_e2e.py knurling | tee -a gpurun_out/r2b_e2e_variants.jsonl
echo "== launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2b_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong --no-cold > gpurun_out/r2b_bench_under_ncu.log 2>&1; echo "launches rc=$?"
echo "== ncu full"
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
cap r2b_tile_gearlike     gearlike 512 jit_kernel 10
cap r2b_extract_gearlike  gearlike 512 extract_kernel 3
cap r2b_tile_blobby       blobby 768 jit_kernel 10
ls -la gpurun_out/r2b*.ncu-rep

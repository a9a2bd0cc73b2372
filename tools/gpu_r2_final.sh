#!/bin/bash
# Round 2, final single-GPU evidence call: whole GPU suite, the bench line (both arms), launch list and ncu captures
# of the kernels the bench runs.
mkdir -p gpurun_out
echo "== pytest gpu (all)"
timeout 1200 python -m pytest tests -m "gpu" -q --timeout 600 > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2f_pytest_gpu.log
echo "== bench"
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
tail -2 gpurun_out/r2f_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
    r = json.loads(open('gpurun_out/r2f_bench_ref.json').read().strip().splitlines()[-1])
    print('value %.0f ms %.4f e2e %.0f (%.3f ms) cold %s roofline %.3f traffic %s cpu %s | ref %.1f Mpts/s %.1f ms' % (
        d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('cold_ms'), d['roofline']['frac'], d['roofline']['traffic'],
        d.get('cpu_baseline', {}).get('value'), r['value'], r['ms_per_step']))
    print(d['kernel_ms'])
except Exception as e:
    print('bench parse failed', e)
PY
echo "== refine grid A/B"
for b in 8 12 16; do SDFB200_REFINE_BLOCKS=$b PROBE_NO_STREAM=1 timeout 60 python tools/gpu_probe.py gearlike,knurling exact 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('refine blocks $b', d['model'], 'refine_ms', d['refine_ms'], 'fused_ms', d['fused_ms'])"; done | tee gpurun_out/r2f_refine_ab.txt
echo "== launch list"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2f_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong --no-cold > gpurun_out/r2f_bench_under_ncu.log 2>&1; echo "launches rc=$?"
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
  SDFB200_JIT=2 SDFB200_GRAPH=0 timeout 120 ncu --set full --clock-control none --import-source on \
    -k regex:$4 -s $5 -c 1 -f -o gpurun_out/$1 python gpurun_out/_one.py $2 $3 > gpurun_out/$1.log 2>&1
  echo "$1 rc=$? $(tail -1 gpurun_out/$1.log)"
}
cap r2f_prepass_gearlike   gearlike 512 jit_kernel 9
cap r2f_tile_gearlike      gearlike 512 jit_kernel 10
cap r2f_refine_gearlike    gearlike 512 jit_kernel 11
cap r2f_classify_gearlike  gearlike 512 tile_classify 3
cap r2f_extract_gearlike   gearlike 512 extract_kernel 3
ls -la gpurun_out/r2f*.ncu-rep

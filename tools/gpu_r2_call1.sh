#!/bin/bash
# Round 2, first GPU call: correctness of the new kernels (sign-mask tile kernel, two-phase extraction, 16-lane
# pre-pass), the bench line, launch list and ncu captures of the kernels the bench runs.
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
echo "== skimage attempt"
( timeout 60 python -m pip install scikit-image 2>&1 | tail -3; python -c "import skimage; print('skimage', skimage.__version__)" 2>&1 | tail -1;
  timeout 60 python -m pip install --no-index --find-links /opt/wheelhouse scikit-image 2>&1 | tail -2 ) > gpurun_out/r2_skimage_attempt.log 2>&1
tail -4 gpurun_out/r2_skimage_attempt.log
timeout 300 python tests/golden/make_mc_golden.py > gpurun_out/r2_mc_golden.log 2>&1; echo "mc golden rc=$?"; tail -1 gpurun_out/r2_mc_golden.log
[ -f tests/golden/mc_skimage.npz ] && cp tests/golden/mc_skimage.npz gpurun_out/
echo "== smoke"
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
echo "== pytest gpu (not fullsize)"
timeout 600 python -m pytest tests -m "gpu and not fullsize" -x -q --timeout 300 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest_gpu.log
echo "== pytest fullsize"
timeout 900 python -m pytest tests -m "fullsize" -q --timeout 600 > gpurun_out/r2_pytest_fullsize.log 2>&1; echo "fullsize rc=$?"
tail -8 gpurun_out/r2_pytest_fullsize.log
echo "== probe"
PROBE_NO_STREAM=1 timeout 300 python tools/gpu_probe.py sphere,csg,gearlike,blobby,knurling exact > gpurun_out/r2_probe.jsonl 2> gpurun_out/r2_probe.err; echo "probe rc=$?"
cat gpurun_out/r2_probe.jsonl | cut -c1-600
echo "== probe spill_ok=96"
SDFB200_JIT_SPILL_OK=96 PROBE_NO_STREAM=1 timeout 300 python tools/gpu_probe.py blobby,knurling,csg exact > gpurun_out/r2_probe_spill96.jsonl 2> gpurun_out/r2_probe_spill96.err; echo "probe rc=$?"
cat gpurun_out/r2_probe_spill96.jsonl | cut -c1-600
echo "== bench"
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
tail -2 gpurun_out/r2_bench1.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2_bench1.json').read().strip().splitlines()[-1])
    print('value %.0f ms %.4f e2e %.0f (%.3f ms) kernels %s roofline %.0f %.3f' % (
        d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['kernel_ms'], d['roofline']['achieved'], d['roofline']['frac']))
except Exception as e:
    print('bench parse failed', e)
PY
echo "== launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1; echo "launches rc=$?"
grep -c . gpurun_out/r2_launches.csv
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
  SDFB200_JIT=2 timeout 200 ncu --set full --clock-control none --import-source on \
    -k regex:$4 -s $5 -c 1 -f -o gpurun_out/$1 python gpurun_out/_one.py $2 $3 > gpurun_out/$1.log 2>&1
  echo "$1 rc=$? $(tail -1 gpurun_out/$1.log)"
}
# per generate() the jit_kernel launches are: skip, tile, refine -> in the 3rd call after the count-only call (4th generate): 9, 10, 11
cap r2_prepass_gearlike  gearlike 512 jit_kernel 9
cap r2_tile_gearlike     gearlike 512 jit_kernel 10
cap r2_refine_gearlike   gearlike 512 jit_kernel 11
cap r2_extract_gearlike  gearlike 512 extract_kernel 3
cap r2_tile_knurling     knurling 1024 jit_kernel 10
ls -la gpurun_out/*.ncu-rep
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tail -1

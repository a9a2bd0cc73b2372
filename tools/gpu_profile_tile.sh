#!/bin/bash
# Round-2 opener (one GPU, ~3 min): ncu --set full source-level captures of the specialised tile kernel in
# both sample groupings on a cheap tree (sphere 256^3: the grouping LOSES there) and on gearlike 512^3 (it
# wins), plus the extraction and refinement kernels.  Reports land in gpurun_out/, read them here with
#   python tools/summarize_ncu.py gpurun_out/<name>.ncu-rep profiles/r2_<name>.txt "<note>"
# jit_kernel launches per generate(): skip pre-pass, tile (148 x 1024), refine; -k regex + launch-skip
# select the tile kernel of a warm call.
mkdir -p gpurun_out
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
cap() {   # name model grid share group kernel-regex launches-to-skip
  SDFB200_JIT=2 SDFB200_JIT_SHARE=$4 SDFB200_JIT_N=$5 timeout 150 ncu --set full --clock-control none --import-source on \
    -k regex:$6 -s $7 -c 1 -f -o gpurun_out/$1 python gpurun_out/_one.py $2 $3 > gpurun_out/$1.log 2>&1
  echo "$1 rc=$? $(tail -1 gpurun_out/$1.log)"
}
# per generate() the jit_kernel launches are: skip, tile, refine  -> the tile kernel of the 3rd call is #7 (0-based skip 7)
cap r2_tile_sphere256_flat   sphere 256 -1 4  jit_kernel 7
cap r2_tile_sphere256_col11  sphere 256  2 11 jit_kernel 7
cap r2_tile_gearlike_flat    gearlike 512 -1 4  jit_kernel 7
cap r2_tile_gearlike_col11   gearlike 512  2 11 jit_kernel 7
cap r2_refine_gearlike       gearlike 512  2 11 jit_kernel 8
cap r2_extract_gearlike      gearlike 512  2 11 extract_kernel 2
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
mkdir -p gpurun_out
for j in 0 2; do SDFB200_JIT=$j timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1; done
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench3.json'))
print('value %.0f Mpts/s  ms %.3f  e2e %.0f Mpts/s  e2e_ms %.3f  pipe_ms %.3f prepass %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['kernel_ms']['fused_tile_kernel'], d['kernel_ms']['skip_prepass+setup']))
PY
tail -2 gpurun_out/bench3.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:jit_kernel -s 10 -c 1 -o gpurun_out/prof_tileA_r1b python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tileA.log 2>&1
ls -la gpurun_out/prof_tileA_r1b.ncu-rep

#!/bin/bash
mkdir -p gpurun_out
for j in 0 2; do SDFB200_JIT=$j timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1; done
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench3.json'))
print('value %.0f Mpts/s  ms %.3f  e2e %.0f Mpts/s  e2e_ms %.3f  pipe_ms %.3f prepass %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['kernel_ms']['fused_tile_kernel'], d['kernel_ms']['skip_prepass+setup']))
PY
tail -2 gpurun_out/bench3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/launches_r1d.csv') if not l.startswith('==')))
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); ig=h.index('Grid Size'); ib=h.index('Block Size')
for r in rows[-6:]:
    print(r[ik][:50].ljust(50), r[ig], r[ib], r[iv])
PY
timeout 300 python tools/gpu_probe.py knurling,blobby exact > gpurun_out/probe6.log 2>&1; grep -h "model" gpurun_out/probe6.log | cut -c1-400

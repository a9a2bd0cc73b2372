#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/gpu_probe.py gearlike,knurling,sphere f32,mixed,exact,f64 > gpurun_out/probe3.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/probe3.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    if 'model' in d: print(d['model'],d['prec'],'fused_ms',d['fused_ms'],'eval Gpts/s',d['gpts_eval_per_s'],'covered',d['gpts_covered_per_s'],'Mtri/s',d['mtris_per_s'],'refined',d['refined'],'tri',d['tri'])
    else: print(d)
PY
for z in 1 0; do
  SDFB200_ZEROCOPY=$z python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision exact > gpurun_out/bench_z$z.json 2> gpurun_out/bench_z$z.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_z$z.json'))
print('zerocopy=$z value %.0f Mpts/s  ms %.3f  e2e %.0f Mpts/s  e2e_ms %.3f  fused_ms %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['kernel_ms']['fused_tile_kernel']))
PY
  tail -2 gpurun_out/bench_z$z.err
done

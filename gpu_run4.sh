#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/gpu_probe.py > gpurun_out/probe2.log 2>&1
grep -v "^+" gpurun_out/probe2.log | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    if 'model' in d: print(d['model'],d['prec'],'fused_ms',d['fused_ms'],'eval Gpts/s',d['gpts_eval_per_s'],'covered',d['gpts_covered_per_s'],'Mtri/s',d['mtris_per_s'],'refined',d['refined'],'tri',d['tri'])
    else: print(d)
"

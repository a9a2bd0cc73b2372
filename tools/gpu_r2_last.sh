#!/bin/bash
# last check of the round: whole GPU suite + the bench line at HEAD
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m "gpu" -q -x --timeout 600 > gpurun_out/r2g_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2g_pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-strong > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2g_bench.json').read().strip().splitlines()[-1])
print('value %.0f ms %.4f e2e %.0f (%.3f ms) roofline %.3f traffic %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic']))
print(d['kernel_ms'])
PY

#!/bin/bash
# One gpurun call: A/B timing of the speed-only mechanisms, the whole GPU test-suite, the bench line,
# and a memcheck of a small generate() on the specialised kernels.  Every step has its own timeout and
# writes under gpurun_out/ as it goes.
mkdir -p gpurun_out
echo "== ab"; timeout 170 python tools/gpu_ab.py --models gearlike,knurling,csg,blobby,sphere > gpurun_out/ab.jsonl 2> gpurun_out/ab.err; echo "ab rc=$?"
tail -2 gpurun_out/ab.err
echo "== pytest"; timeout 260 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 120 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
tail -2 gpurun_out/bench_final.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
    print('value %.0f ms %.4f e2e %.0f (%.3f ms) kernels %s roofline %.0f %.3f grouping %s' % (
        d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['kernel_ms'], d['roofline']['achieved'],
        d['roofline']['frac'], d['config'].get('sample_grouping')))
except Exception as e:
    print('bench parse failed', e)
for l in open('gpurun_out/ab.jsonl'):
    d = json.loads(l)
    if d['kind'] == 'grouping':
        print('%-9s %-6s ax %2d N %2d tile %.4f refine %.4f extract %.4f pipe %.4f tri %d ck %x' % (d['model'], d['variant'], d['share_axis'], d['group'], d['tile_eval_ms'], d['refine_ms'], d['extract_ms'], d['pipeline_ms'], d['triangles'], d['checksum']))
    else:
        print('%-9s pipe %d e2e %.3f ms (min %.3f) tri %d ck %x' % (d['model'], d['pipe_rounds'], d['e2e_ms_median'], d['e2e_ms_min'], d['triangles'], d['checksum']))
PY
echo "== memcheck"; SDFB200_JIT=2 timeout 90 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/memcheck.log

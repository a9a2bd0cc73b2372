#!/bin/bash
# Short follow-up call: A/B of the sample grouping after the thin-tile skip, the tests that depend on
# batch classification and on the grouping / pipelined output, the bench line.
mkdir -p gpurun_out
echo "== ab"
timeout 45 python tools/gpu_ab.py --quick --models csg,sphere,gearlike,blobby > gpurun_out/ab2.jsonl 2> gpurun_out/ab2.err; echo "ab rc=$?"
timeout 20 python tools/gpu_ab.py --quick --models sphere --grid 256 >> gpurun_out/ab2.jsonl 2>> gpurun_out/ab2.err; echo "ab256 rc=$?"
tail -2 gpurun_out/ab2.err
echo "== pytest"
timeout 110 python -m pytest tests -m gpu -q --timeout 100 -k "edge_cases or sparse_off or auto_bounds or exact_mode or grouping or pipelined or (generate_matches_oracle and (sphere or csg))" > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu2.log
echo "== bench"
timeout 70 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo "bench rc=$?"
tail -2 gpurun_out/bench_final2.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_final2.json').read().strip().splitlines()[-1])
    print('value %.0f ms %.4f e2e %.0f (%.3f ms) kernels %s roofline %.0f %.3f' % (
        d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['kernel_ms'], d['roofline']['achieved'], d['roofline']['frac']))
except Exception as e:
    print('bench parse failed', e)
for l in open('gpurun_out/ab2.jsonl'):
    d = json.loads(l)
    if d['kind'] == 'grouping':
        print('%-9s %4d %-6s ax %2d N %2d tile %.4f (min %.4f) refine %.4f extract %.4f pipe %.4f tri %d ck %x' % (d['model'], d['grid'], d['variant'], d['share_axis'], d['group'], d['tile_eval_ms'], d['tile_eval_min_ms'], d['refine_ms'], d['extract_ms'], d['pipeline_ms'], d['triangles'], d['checksum']))
    else:
        print('%-9s %4d pipe %d e2e %.3f ms (min %.3f) tri %d ck %x' % (d['model'], d['grid'], d['pipe_rounds'], d['e2e_ms_median'], d['e2e_ms_min'], d['triangles'], d['checksum']))
PY

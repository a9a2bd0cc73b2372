#!/bin/bash
# N-GPU call: windows of the Z-major batch order == single GPU (both gathers, both partitions), then the bench line
# at N GPUs (weak series + the strong 1024^3 side series).
N=$1
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/gpu_multi_check.py 2>&1 | grep -v "^W0\|^\*\*\*" | tail -24 | tee gpurun_out/r2_multi_check_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -3 gpurun_out/r2_bench_n$N.err; python - <<PY
import json
d = json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
print('N=%d value %.0f Mpts/s ms %.4f e2e %.0f (%.3f ms) tri %d grid %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['triangles'], d['config']['workload']))
print('kernels', d['kernel_ms'])
print('strong', d['strong_1024'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-rebalance --no-cpu-baseline > gpurun_out/r2_bench_n${N}_norebalance.json 2> gpurun_out/r2_bench_n${N}_norebalance.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_bench_n${N}_norebalance.json').read().strip().splitlines()[-1])
    print('no re-plan: N=%d value %.0f ms %.4f strong %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['strong_1024']['ms_per_step'] if d.get('strong_1024') else None))
except Exception as e:
    print('parse failed', e)
PY

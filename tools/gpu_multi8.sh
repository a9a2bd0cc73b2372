#!/bin/bash
N=$1
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/gpu_multi_check.py 2>&1 | grep -E "CHECK|identical=False" | tail -3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --model knurling --grid 1024 > gpurun_out/bench_knurl1024_n$N.json 2> gpurun_out/bench_k_n$N.err
timeout 100 python bench.py --steps 10 --warmup 3 --model knurling --grid 1024 --no-cpu-baseline > gpurun_out/bench_knurl1024_n1.json 2>/dev/null
for f in bench_n$N bench_knurl1024_n$N bench_knurl1024_n1; do python -c "
import sys,json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
print('$f N=%d value %.0f Mpts/s ms %.3f e2e %.0f (%.3f ms) tri %d %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['triangles'], d['config']['workload'][:40]))
"; done
tail -2 gpurun_out/bench_n$N.err

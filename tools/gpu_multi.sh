#!/bin/bash
N=$1
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/gpu_multi_check.py 2>&1 | grep -v "^W0\|^\*\*\*" | tail -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -3 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=%d value %.0f Mpts/s ms %.3f e2e %.0f tri %d grid %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['triangles'], d['config']['workload']))
"

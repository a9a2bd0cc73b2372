#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/gpu_probe.py > gpurun_out/probe.log 2>&1
tail -30 gpurun_out/probe.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
cat gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err

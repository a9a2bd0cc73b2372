#!/bin/bash
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_tile -s 3 -c 1 -o gpurun_out/prof_fused_r1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fused.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_stream_f32 -s 6 -c 1 -o gpurun_out/prof_stream_r1 python tools/gpu_probe.py gearlike f32 > gpurun_out/ncu_stream.log 2>&1
ls -la gpurun_out

#!/bin/bash
mkdir -p gpurun_out
# bench (N=1), default flags; and the reference arm
timeout 400 python bench.py > gpurun_out/BENCH_self_r1.json 2> gpurun_out/bench_self.err; tail -2 gpurun_out/bench_self.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/BENCH_ref_r1.json 2>/dev/null
# launch list of the same command
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
# full captures: tile kernel (2nd jit launch of a step), refine (3rd), extract, and the streamed evaluator
timeout 400 ncu --set full --clock-control none --import-source on -k regex:jit_kernel -s 10 -c 1 -o gpurun_out/r1_tile_eval python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:jit_kernel -s 11 -c 1 -o gpurun_out/r1_refine python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:extract_kernel -s 3 -c 1 -o gpurun_out/r1_extract python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:jit_kernel -s 8 -c 1 -o gpurun_out/r1_stream_fast_gearlike python tools/gpu_probe.py gearlike exact > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:jit_kernel -s 8 -c 1 -o gpurun_out/r1_stream_fast_sphere python tools/gpu_probe.py sphere exact > /dev/null 2>&1
timeout 300 python tools/gpu_probe.py > gpurun_out/r1_probe_all.log 2>&1
ls -la gpurun_out | tail -15
cat gpurun_out/BENCH_self_r1.json | cut -c1-600

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -4
timeout 400 python bench.py > gpurun_out/BENCH_self_r1.json 2> gpurun_out/bench_self.err; tail -2 gpurun_out/bench_self.err
python -c "
import json; d=json.load(open('gpurun_out/BENCH_self_r1.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'kernels',d['kernel_ms'],'roofline',d['roofline']['achieved'],d['roofline']['frac'],d['roofline']['traffic'],'clocks',d['clocks'])"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:jit_kernel -s 24 -c 1 -o gpurun_out/r1_stream_fast_gearlike python tools/gpu_probe.py gearlike exact > /dev/null 2>&1
ls -la gpurun_out/r1_stream_fast_gearlike.ncu-rep

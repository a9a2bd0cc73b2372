#!/bin/bash
mkdir -p gpurun_out
for j in 0 2; do SDFB200_JIT=$j timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1; done
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 --durations=8 2>&1 | tail -25

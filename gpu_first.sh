#!/bin/bash
# first GPU contact: smoke, parity tests, a timing probe
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20
python -m pytest tests -m gpu -x -q 2>&1 | tail -30

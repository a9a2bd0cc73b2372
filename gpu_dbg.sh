#!/bin/bash
timeout 300 python -m pytest tests -m gpu -x -q --timeout 300 -k "test_generate_matches_oracle and jit and mixed and gearlike" 2>&1 | grep -E "^E " | head -12

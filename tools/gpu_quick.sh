#!/bin/bash
# parity sweep (fixed + generic selector instantiations) and timing of the bench workloads
mkdir -p gpurun_out
echo "=== fixed selectors"; timeout 600 python tools/gpu_check.py 2>&1 | tail -22
echo "=== generic selectors"; AMPE_B200_GENERIC=1 timeout 600 python tools/gpu_check.py 2>&1 | tail -22
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5

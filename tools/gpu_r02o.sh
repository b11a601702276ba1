#!/bin/bash
# Round 2, call O: the suite with the final library (no -x: every failure listed)
mkdir -p gpurun_out
timeout -k 5 1800 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r02o_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02o_pytest_gpu.log
tail -16 gpurun_out/r02o_pytest_gpu.log

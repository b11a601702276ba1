#!/bin/bash
# round 2, call Z: grain diagnostics on the device vs the restatement; the TwoGrainsQuadratic deck's grain-volume lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_grains.py tests/test_regression_decks.py -q -m gpu -s -k "grain or two_grains" > gpurun_out/r02z_pytest_grains.log 2>&1
grep -E "grain volumes over|passed|failed|Error" gpurun_out/r02z_pytest_grains.log | cut -c1-1500

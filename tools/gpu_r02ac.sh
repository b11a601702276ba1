#!/bin/bash
# round 2, call AC: the 3D versions of the reference's decks on the device
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_regression_decks.py -q -m gpu -s -k "3d_gpu" > gpurun_out/r02ac_pytest_decks3d.log 2>&1
grep -E "3D:|passed|failed|Error|assert |^E  " gpurun_out/r02ac_pytest_decks3d.log | cut -c1-700 | head -40

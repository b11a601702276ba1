#!/bin/bash
# round 2, call AF: SolidifyQuaternions deck on the device
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_regression_decks.py -q -m gpu -s -k "solidify_quaternions" > gpurun_out/r02af_pytest_solidify.log 2>&1
grep -E "SolidifyQuaternions:|passed|failed|Error|assert |^E  " gpurun_out/r02af_pytest_solidify.log | cut -c1-900 | head -20

#!/bin/bash
# round 2, call AF: SolidifyQuaternions deck on the device
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_regression_decks.py -q -m gpu -s -k "solidify_quaternions_deck_3d" > gpurun_out/r02af_pytest_solidify3d.log 2>&1
grep -E "SolidifyQuaternions 3D:|passed|failed|Error|assert |^E  " gpurun_out/r02af_pytest_solidify3d.log | cut -c1-900 | head -20

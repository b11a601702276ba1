#!/bin/bash
# round 2, call AA: the FourCorners deck (qlen 4 in 2D, evolving quaternions, grain volumes) on the device
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_regression_decks.py -q -m gpu -s -k "four_corners" > gpurun_out/r02aa_pytest_four_corners.log 2>&1
grep -E "four corners:|passed|failed|Error|assert" gpurun_out/r02aa_pytest_four_corners.log | cut -c1-1800

#!/bin/bash
# round 2, call AD: TwoGrainsQuadratic 2D deck on the device
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_regression_decks.py -q -m gpu -s -k "quadratic_deck_2d" > gpurun_out/r02ad_pytest_tgq2d.log 2>&1
grep -E "2D:|passed|failed|Error|assert |^E  " gpurun_out/r02ad_pytest_tgq2d.log | cut -c1-900 | head -20

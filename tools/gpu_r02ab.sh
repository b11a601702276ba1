#!/bin/bash
# round 2, call AB: KKS composition flux with the CALPHAD free energy -- parity cases and the KKScomposition deck
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_regression_decks.py -q -m gpu -s -k "kks_flux_calphad or kks_composition" > gpurun_out/r02ab_pytest_kks_calphad.log 2>&1
grep -E "KKScomposition:|passed|failed|Error|assert " gpurun_out/r02ab_pytest_kks_calphad.log | cut -c1-600

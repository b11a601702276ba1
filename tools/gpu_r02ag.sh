#!/bin/bash
# round 2, call AG: EBS flux with the quadratic free energy (Tbased diffusion): parity and the OneGrainQuadratic decks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_regression_decks.py -q -m gpu -s -k "ebs_flux_with_quadratic or one_grain_quadratic" > gpurun_out/r02ag_pytest_ebs_quadratic.log 2>&1
grep -E "OneGrainQuadratic|passed|failed|Error|assert |^E  " gpurun_out/r02ag_pytest_ebs_quadratic.log | cut -c1-700 | head -30

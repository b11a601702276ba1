#!/bin/bash
# Round 2, call G (one GPU): BASELINE-size parity, the process-level halo tests with the in-kernel wait, the suite
mkdir -p gpurun_out
nproc; free -g | head -2
timeout -k 5 1500 python -m pytest tests/test_gpu_parity_fullsize.py -m gpu -q -s --durations=5 > gpurun_out/r02g_pytest_fullsize.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest_fullsize.log
grep -E "^(dendrite2d|auni2d|gg3d_hbsm|auni3d) |passed|failed|rc=" gpurun_out/r02g_pytest_fullsize.log | cut -c1-900
timeout -k 5 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity_fullsize.py > gpurun_out/r02g_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest_gpu.log
tail -6 gpurun_out/r02g_pytest_gpu.log

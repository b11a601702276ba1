#!/bin/bash
# Round 2, call F: zero-slope / ramp / DeltaT parity and the reference's regression decks on the device, then the suite
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests/test_regression_decks.py "tests/test_gpu_parity.py::test_zero_slope_boundaries_and_temperature_ramp" -m gpu -q --durations=8 > gpurun_out/r02f_pytest_decks.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest_decks.log
tail -30 gpurun_out/r02f_pytest_decks.log
timeout -k 5 900 python -m pytest tests -m gpu -q --deselect tests/test_regression_decks.py > gpurun_out/r02f_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest_gpu.log
tail -6 gpurun_out/r02f_pytest_gpu.log

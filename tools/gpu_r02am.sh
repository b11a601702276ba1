#!/bin/bash
# final check of the host library built from the last commit (implicit integrator option bits, NetCDF-4 reader linked in): the device
# tests of the implicit integrator, the deck program, one preconditioned regression deck
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_widening_z_implicit.py tests/test_run_deck.py "tests/test_regression_decks.py::test_single_grain_auni_deck_gpu" -q -m gpu -p no:cacheprovider > gpurun_out/r02am_pytest_final_host.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02am_pytest_final_host.log
tail -n 4 gpurun_out/r02am_pytest_final_host.log

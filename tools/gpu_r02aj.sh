#!/bin/bash
# last GPU call of round 2 (5 GPU-minutes left): the deck program on the device, the NetCDF-4 reader and the deck reader on the box
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_run_deck.py tests/test_initial_conditions_netcdf4.py tests/test_input_deck.py -q -p no:cacheprovider > gpurun_out/r02aj_pytest_run_deck.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02aj_pytest_run_deck.log
tail -n 5 gpurun_out/r02aj_pytest_run_deck.log

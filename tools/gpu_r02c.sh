#!/bin/bash
# Round 2, call C (one GPU): the GPU suite with the arbiter-based parity check and the process-level halo tests
# (two / three ranks sharing the GPU), then the default bench line.
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest_gpu.log
tail -40 gpurun_out/r02c_pytest_gpu.log
timeout -k 5 300 python bench.py > gpurun_out/r02c_bench_default.json 2> gpurun_out/r02c_bench_default.err
cut -c1-1500 gpurun_out/r02c_bench_default.json; tail -3 gpurun_out/r02c_bench_default.err

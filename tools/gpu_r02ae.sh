#!/bin/bash
# round 2, call AE: the whole GPU suite with the final library + smoke + the default bench line
mkdir -p gpurun_out
timeout -k 5 2400 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r02ae_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ae_pytest_gpu.log
tail -18 gpurun_out/r02ae_pytest_gpu.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r02ae_smoke.log 2>&1; tail -1 gpurun_out/r02ae_smoke.log
timeout -k 5 400 python bench.py > gpurun_out/r02ae_bench_default.json 2> gpurun_out/r02ae_bench.err; tail -c 600 gpurun_out/r02ae_bench_default.json

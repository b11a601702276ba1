#!/bin/bash
# Round 2, call M: the suite with the final library (decks preconditioned, multigrid with zero-slope boundaries), smoke, default bench
mkdir -p gpurun_out
timeout -k 5 1800 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/r02m_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m_pytest_gpu.log
tail -18 gpurun_out/r02m_pytest_gpu.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02m_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02m_smoke.log; tail -3 gpurun_out/r02m_smoke.log | cut -c1-400
timeout -k 5 400 python bench.py > gpurun_out/r02m_bench_default.json 2> gpurun_out/r02m_bench_default.err; cut -c1-700 gpurun_out/r02m_bench_default.json; tail -2 gpurun_out/r02m_bench_default.err
timeout -k 5 400 python bench.py --impl reference > gpurun_out/r02m_bench_reference.json 2>> gpurun_out/r02m_bench_default.err; cut -c1-300 gpurun_out/r02m_bench_reference.json

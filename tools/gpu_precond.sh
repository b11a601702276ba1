#!/bin/bash
# First GPU call for the block preconditioners: the opt-in parity tests, the V-cycle
# timing against the HBM roofline, the ncu launch list and one full capture of the smoother.
mkdir -p gpurun_out
AMPE_B200_RUN_EXPERIMENTS=1 timeout -k 5 400 python -m pytest tests/test_gpu_widening_zzz_precond.py -q -m gpu \
  > gpurun_out/pytest_precond.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_precond.log
tail -15 gpurun_out/pytest_precond.log
timeout -k 5 300 python tools/bench_precond.py > gpurun_out/bench_precond.jsonl 2> gpurun_out/bench_precond.err
timeout -k 5 300 python tools/bench_precond.py --quat >> gpurun_out/bench_precond.jsonl 2>> gpurun_out/bench_precond.err
AMPE_B200_MG_TAIL=0 timeout -k 5 300 python tools/bench_precond.py >> gpurun_out/bench_precond.jsonl 2>> gpurun_out/bench_precond.err
AMPE_B200_MG_FUSED=0 timeout -k 5 300 python tools/bench_precond.py >> gpurun_out/bench_precond.jsonl 2>> gpurun_out/bench_precond.err
AMPE_B200_MG_GRAPH=1 timeout -k 5 300 python tools/bench_precond.py --stream >> gpurun_out/bench_precond.jsonl 2>> gpurun_out/bench_precond.err
cat gpurun_out/bench_precond.jsonl; tail -3 gpurun_out/bench_precond.err
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_precond.csv \
  python tools/bench_precond.py --cases 2d:2048x2048 --cycles 2 --reps 1 > gpurun_out/ncu_launch_precond.log 2>&1
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:'mg_rb_fused|mg_smooth_rb|mg_residual|mg_prolong' -s 4 -c 3 -f \
  -o gpurun_out/prof_precond python tools/bench_precond.py --cases 3d:256x256x256 --cycles 2 --reps 1 > gpurun_out/ncu_full_precond.log 2>&1
ls -la gpurun_out | tail -8

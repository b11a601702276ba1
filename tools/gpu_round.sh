#!/bin/bash
# One gpurun call: GPU parity tests, bench lines for every workload, ncu launch lists + one full
# capture of the dominant kernels.  Every step has its own timeout.
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for w in dendrite2d auni3d; do
  timeout -k 5 240 python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  tail -c 1200 gpurun_out/bench_$w.json
done
for w in auni2d gg3d_hbsm pfhub1a; do
  timeout -k 5 240 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  tail -c 600 gpurun_out/bench_$w.json
done
timeout -k 5 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2>&1
for w in dendrite2d auni3d; do
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$w.csv \
  python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_$w.log 2>&1
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:'rhs_|kks_' -s 6 -c 2 -f -o gpurun_out/prof_$w \
  python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_$w.log 2>&1
done
ls -la gpurun_out | tail -20

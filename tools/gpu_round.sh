#!/bin/bash
# One gpurun call: GPU parity tests, bench lines for every workload, ncu launch list + one full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for w in dendrite2d auni2d gg3d_hbsm auni3d pfhub1a; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  tail -c 1500 gpurun_out/bench_$w.json
done
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2>&1
for w in dendrite2d auni3d; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$w.csv \
  python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_$w.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rhs_ -s 3 -c 2 -f -o gpurun_out/prof_$w \
  python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_$w.log 2>&1
done
kill $SMI
ls -la gpurun_out

#!/bin/bash
# round 2, call V1 (one GPU): final library -- whole GPU suite, smoke, ncu launch lists and full captures
mkdir -p gpurun_out
timeout -k 5 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02v_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02v_pytest_gpu.log
tail -14 gpurun_out/r02v_pytest_gpu.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r02v_smoke.log 2>&1; tail -3 gpurun_out/r02v_smoke.log
for w in auni3d dendrite2d; do
  timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02v_launches_$w.csv \
    python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ncu_launch_$w.log 2>&1
done
for w in auni3d dendrite2d auni2d gg3d_hbsm; do
  timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:'rhs_|kks_' -s 8 -c 2 -f -o gpurun_out/prof_r02v_$w \
    python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ncu_full_r02v_$w.log 2>&1
done
ls -la gpurun_out | grep r02v | tail -20

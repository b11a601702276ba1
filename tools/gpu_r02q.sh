#!/bin/bash
# round 2, call Q: driving force staged with the plane (no global read in the cell phase) -- parity, bench, fresh ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -4
timeout -k 5 300 python bench.py --workload auni3d --no-e2e --no-cpu-baseline > gpurun_out/r02q_bench_auni3d.json 2> gpurun_out/r02q_bench.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02q_bench_auni3d.json') if l.startswith('{')][-1])
print('auni3d ms', d['ms_per_step'], [ (k['kernel'][:20], round(k['ms'],3)) for k in d['roofline']['kernels']])
PY
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:'rhs_march' -s 4 -c 1 -f -o gpurun_out/prof_r02q_auni3d \
    python bench.py --workload auni3d --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ncu_full_r02q_auni3d.log 2>&1
ls -la gpurun_out | grep r02q

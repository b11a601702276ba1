#!/bin/bash
# round 2, call AI: the default GPU suite and the default bench line with the final library of the round
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r02ai_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ai_pytest_gpu.log
tail -14 gpurun_out/r02ai_pytest_gpu.log
timeout -k 5 300 python bench.py --no-cpu-baseline > gpurun_out/r02ai_bench_default.json 2> gpurun_out/r02ai_bench.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02ai_bench_default.json') if l.startswith('{')][-1])
print('auni3d ms', d['ms_per_step'], 'GCUPS', d['value'], 'e2e', d['e2e'] and d['e2e']['value'], [ (k['kernel'][:5], round(k['ms'],3)) for k in d['roofline']['kernels']])
PY

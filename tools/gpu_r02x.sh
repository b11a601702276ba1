#!/bin/bash
# round 2, call X: marching-kernel column heights 7 and 6 (smaller shared-memory carve-out / three blocks per SM)
mkdir -p gpurun_out
for v in default my7 my6 default; do
  if [ $v = default ]; then unset AMPE_B200_LIB; else export AMPE_B200_LIB=$PWD/variants/lib_$v.so; fi
  if [ $v != default ]; then timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "rhs_matches_oracle and (auni3d or gg3d)" 2>&1 | tail -1; fi
  for w in auni3d gg3d_hbsm; do
  timeout -k 5 300 python bench.py --workload $w --no-e2e --no-cpu-baseline --no-extras --steps 10 > gpurun_out/r02x_x.json 2> gpurun_out/r02x_x.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02x_x.json') if l.startswith('{')][-1])
print('$v', '$w', 'ms', round(d['ms_per_step'],3), [ (k['kernel'][:5], round(k['ms'],3)) for k in d['roofline']['kernels']])
PY
  done
done 2>&1 | tee gpurun_out/r02x_ab.log

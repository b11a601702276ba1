#!/bin/bash
# round 2, call R: one copy of the face code for edge + owner warps (A/B), explicit register caps for the nine-warp block
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "auni3d or gg3d or anisotropic_3d" 2>&1 | tail -3
for v in default nounified nreg104 nreg112 default; do
  if [ $v = default ]; then unset AMPE_B200_LIB; else export AMPE_B200_LIB=$PWD/variants/lib_$v.so; fi
  timeout -k 5 300 python bench.py --workload auni3d --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r02r_$v.json 2> gpurun_out/r02r_$v.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02r_$v.json') if l.startswith('{')][-1])
print('$v', 'auni3d ms', round(d['ms_per_step'],3), [ (k['kernel'][:10], round(k['ms'],3)) for k in d['roofline']['kernels']])
PY
done 2>&1 | tee gpurun_out/r02r_ab.log

#!/bin/bash
# Round-2 first call: A/B of the split 3D launches (AMPE_B200_SPLIT3D=1, rhs_march.cuh PART 1/2)
# against the fused march kernel on one box, for the register caps prepared as variants:
#   tools/build_variant.sh s34 "-DAMPE_SPLIT_MINB1=3 -DAMPE_SPLIT_MINB2=4"   (etc., build BEFORE gpurun)
# usage (on the GPU box): tools/gpu_split3d.sh [variant ...]
set -u
run() { # label env-assignments...
  local label=$1; shift
  env "$@" timeout -k 5 180 python bench.py --workload auni3d --steps 20 --warmup 5 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$label: ms/step %.3f  GCUPS %.2f  launches/step %d  clocks %s'%(d['ms_per_step'], d['value'], d['gpu_launches']//d['steps'], d['clocks']['sm_mhz']))
    elif 'rror' in l: print(l.strip()[:300])"
}
AMPE_B200_RUN_EXPERIMENTS=1 python -m pytest tests/test_gpu_widening_zz_split3d.py tests/test_gpu_widening_z_implicit.py -q -m gpu 2>&1 | tail -3
for rep in 1 2; do
  run fused X=1
  run split-default AMPE_B200_SPLIT3D=1
  for v in "$@"; do run split-$v AMPE_B200_SPLIT3D=1 AMPE_B200_LIB=$PWD/variants/lib_$v.so; done
done

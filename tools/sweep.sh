#!/bin/bash
# time the bench workloads with each library variant in variants/ (development aid)
# lib_t2_* variants only differ for 2D workloads, lib_t3_* for 3D ones
run() {
  lib=$1; w=$2
  AMPE_B200_LIB=$PWD/$lib timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$lib', '$w', '%.3f ms  %.2f GCUPS  frac %.3f' % (d['ms_per_step'], d['value'], d['roofline']['frac']))
except Exception as e: print('$lib $w failed', e)"
}
for w in dendrite2d auni2d gg3d_hbsm auni3d; do run ampe_b200/libampe_b200.so $w; done
for lib in variants/lib_*.so; do
  case $lib in
    *lib_t2_*) ws="dendrite2d auni2d";;
    *lib_t3_*) ws="gg3d_hbsm auni3d";;
    *) ws="dendrite2d auni2d gg3d_hbsm auni3d";;
  esac
  for w in $ws; do run $lib $w; done
done

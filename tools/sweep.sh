#!/bin/bash
# time the bench workloads with each library variant in variants/ (development aid)
for lib in ampe_b200/libampe_b200.so variants/lib_*.so; do
  for w in dendrite2d gg3d_hbsm auni3d auni2d; do
    AMPE_B200_LIB=$PWD/$lib python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$lib', '$w', '%.3f ms  %.2f GCUPS  frac %.3f' % (d['ms_per_step'], d['value'], d['roofline']['frac']))
except Exception as e: print('$lib $w failed', e)"
  done
done

#!/bin/bash
# A/B of two builds of the library on the same box: usage gpu_ab.sh <variant-name> [workloads...]
V=$1; shift
for rep in 1 2; do
for lib in default $V; do
  for w in "$@"; do
    if [ $lib = default ]; then L=""; else L="AMPE_B200_LIB=$PWD/variants/lib_$lib.so"; fi
    env $L timeout -k 5 120 python bench.py --workload $w --steps 30 --warmup 5 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$lib $w: ms/step %.4f  GCUPS %.2f  clocks %s'%(d['ms_per_step'], d['value'], d['clocks']['sm_mhz']))
    elif 'rror' in l: print(l.strip()[:300])"
  done
done
done

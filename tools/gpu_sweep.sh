#!/bin/bash
# tuning sweep on one box: "variant:workload[,workload]" pairs, default library first
run() { # lib workload
  local lib=$1 w=$2 L=""
  [ $lib != default ] && L="AMPE_B200_LIB=$PWD/variants/lib_$lib.so"
  env $L timeout -k 5 120 python bench.py --workload $w --steps 30 --warmup 5 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$lib $w: ms/step %.4f  GCUPS %.2f'%(d['ms_per_step'], d['value']))
    elif 'rror' in l: print(l.strip()[:300])"
}
for w in dendrite2d auni3d gg3d_hbsm; do run default $w; done
for spec in "$@"; do
  lib=${spec%%:*}; ws=${spec#*:}
  for w in ${ws//,/ }; do run $lib $w; done
done

#!/bin/bash
# GPU parity tests + device-resident timing of every workload (no e2e / CPU legs)
mkdir -p gpurun_out
timeout -k 5 240 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in ${WORKLOADS:-dendrite2d auni2d gg3d_hbsm auni3d}; do
  env "$@" timeout -k 5 120 python bench.py --workload $w --steps 30 --warmup 5 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$w: ms/step %.4f  GCUPS %.2f  frac %.3f nf %s'%(d['ms_per_step'], d['value'], d['roofline']['frac'], d['newton_failures']))
    elif 'rror' in l: print(l.strip()[:300])"
done

#!/bin/bash
# TMA persistent 2D kernel: parity first, then timing against the cp.async tile kernel
mkdir -p gpurun_out
timeout -k 5 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tma or full_size or matches_oracle or ragged" 2>&1 | tail -6
run() { # label, env...
  local label=$1; shift
  for w in dendrite2d auni2d; do
    env "$@" timeout -k 5 90 python bench.py --workload $w --steps 50 --warmup 5 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$label $w: ms/step %.4f  GCUPS %.2f  frac %.3f'%(d['ms_per_step'], d['value'], d['roofline']['frac']))
    elif 'rror' in l: print(l.strip()[:300])"
  done
}
run "tma32x32/512" AMPE_B200_TMA=1
run "tma32x16/256" AMPE_B200_TMA=1 AMPE_B200_LIB=$PWD/variants/lib_tma16.so
run "cp.async tile" A=1

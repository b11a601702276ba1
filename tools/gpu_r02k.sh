#!/bin/bash
# Round 2, call K: where did the marching kernel lose 2.8 % between r02e and r02i?  clamp code | in-kernel wait | both
mkdir -p gpurun_out
line() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        ks=' | '.join('%s %.3f ms'%(k['kernel'][:10],k['ms']) for k in d['roofline']['kernels'])
        print('$1 $2: ms/step %.4f GCUPS %.2f [%s] clocks %s'%(d['ms_per_step'],d['value'],ks,d['clocks']['sm_mhz']))
    elif 'rror' in l: print(l.strip()[:300])"; }
run() { # lib workload
  if [ $1 = default ]; then L="X=1"; else L="AMPE_B200_LIB=$PWD/variants/lib_$1.so"; fi
  env $L timeout -k 5 200 python bench.py --workload $2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extras 2>&1 | tee -a gpurun_out/r02k_ab_$1.jsonl | line $1 $2
}
for rep in 1 2; do
  for lib in default noclamp nowait neither edgewarp; do
    run $lib auni3d
    run $lib dendrite2d
    run $lib gg3d_hbsm
  done
  for lib in t2y8 unrollrows; do
    run $lib dendrite2d
    run $lib auni2d
  done
  run default auni2d
done 2>&1 | tee gpurun_out/r02k_ab.log

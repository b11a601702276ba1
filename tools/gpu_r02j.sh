#!/bin/bash
# Round 2, call J: A/B of the marching-kernel variants (edge warp, 32x4 columns, out-of-line faces) on one box
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
  env $L timeout -k 5 200 python bench.py --workload $2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extras 2>&1 | tee -a gpurun_out/r02j_ab_$1.jsonl | line $1 $2
}
for rep in 1 2; do
  for lib in default edgewarp my4 my4e facecall; do
    run $lib auni3d
    run $lib gg3d_hbsm
  done
  run facecall auni2d
  run default auni2d
done 2>&1 | tee gpurun_out/r02j_ab.log
for v in edgewarp my4e facecall; do
  AMPE_B200_LIB=$PWD/variants/lib_$v.so timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "oracle or golden or ragged or split or two_slab" > gpurun_out/r02j_pytest_$v.log 2>&1
  echo "== parity tests with lib_$v: $(tail -1 gpurun_out/r02j_pytest_$v.log)"; grep FAILED gpurun_out/r02j_pytest_$v.log | head -10
done

#!/bin/bash
# Round 2, call E: GPU suite (default library: Horner/FMA Newton polynomials, hoisted staging arithmetic), then A/B
# of the variants -- straight-line atan (dendrite2d), out-of-line symmetry rotation (auni2d), --fmad=true (all) --
# with the parity tests of each variant library.
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest_gpu.log
tail -5 gpurun_out/r02e_pytest_gpu.log
line() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); n=d.get('newton') or {}
        ks=' | '.join('%s %.3f ms %.1f%%'%(k['kernel'][:10],k['ms'],100*k['frac']) for k in d['roofline']['kernels'])
        print('$1 $2: ms/step %.4f GCUPS %.2f frac %.3f warm %s [%s] clocks %s'%(d['ms_per_step'],d['value'],d['roofline']['frac'],n.get('warm_ms_per_step'),ks,d['clocks']['sm_mhz']))
    elif 'rror' in l: print(l.strip()[:300])"; }
run() { # lib workload
  if [ $1 = default ]; then L="X=1"; else L="AMPE_B200_LIB=$PWD/variants/lib_$1.so"; fi
  env $L timeout -k 5 200 python bench.py --workload $2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tee -a gpurun_out/r02e_ab_$1.jsonl | line $1 $2
}
for rep in 1 2; do
  for w in auni3d gg3d_hbsm auni2d dendrite2d; do run default $w; done
  for w in auni3d gg3d_hbsm auni2d dendrite2d; do run fmad $w; done
  run atanfast dendrite2d
  run noinline auni2d
done 2>&1 | tee gpurun_out/r02e_ab.log
for v in fmad atanfast noinline; do
  AMPE_B200_LIB=$PWD/variants/lib_$v.so timeout -k 5 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_widening.py tests/test_gpu_strategies.py -q -m gpu > gpurun_out/r02e_pytest_$v.log 2>&1
  echo "== parity tests with lib_$v: $(tail -1 gpurun_out/r02e_pytest_$v.log)"; grep FAILED gpurun_out/r02e_pytest_$v.log | head -20
done

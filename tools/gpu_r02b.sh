#!/bin/bash
# Round 2, call B: FP64 issue-rate probe, the GPU suite with the straight-line log / constant-memory tables, A/B of
# the library variants (default | CUDA log in the KKS Newton | --fmad=true), parity of the fmad build, ncu capture.
mkdir -p gpurun_out
tools/probe/fp64_peak > gpurun_out/r02b_fp64_peak.txt 2>&1; tail -4 gpurun_out/r02b_fp64_peak.txt
timeout -k 5 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest_gpu.log
tail -4 gpurun_out/r02b_pytest_gpu.log
line() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); n=d.get('newton') or {}
        ks=' | '.join('%s %.3f ms %.1f%%'%(k['kernel'][:10],k['ms'],100*k['frac']) for k in d['roofline']['kernels'])
        print('$1 $2: ms/step %.4f GCUPS %.2f frac %.3f warm %s cold %s [%s] clocks %s'%(d['ms_per_step'],d['value'],d['roofline']['frac'],n.get('warm_ms_per_step'),n.get('cold_ms_per_step'),ks,d['clocks']['sm_mhz']))
    elif 'rror' in l: print(l.strip()[:300])"; }
for rep in 1 2; do
for lib in default libmlog fmad; do
  for w in auni3d gg3d_hbsm auni2d dendrite2d; do
    if [ $lib = default ]; then L="X=1"; else L="AMPE_B200_LIB=$PWD/variants/lib_$lib.so"; fi
    env $L timeout -k 5 200 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tee -a gpurun_out/r02b_ab_$lib.jsonl | line $lib $w
  done
done
done 2>&1 | tee gpurun_out/r02b_ab.log
AMPE_B200_LIB=$PWD/variants/lib_fmad.so timeout -k 5 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_widening.py -q -m gpu > gpurun_out/r02b_pytest_fmad.log 2>&1
tail -15 gpurun_out/r02b_pytest_fmad.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rhs_|kks_' -s 6 -c 2 -f -o gpurun_out/prof_r02b_auni3d \
  python bench.py --workload auni3d --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ncu_full_r02b_auni3d.log 2>&1
ls -la gpurun_out | tail -8

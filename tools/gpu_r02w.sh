#!/bin/bash
# round 2, call W: KKS pre-pass of chunk j+1 beside the marching kernel of chunk j (A/B against the sequential order)
mkdir -p gpurun_out
run() {  # label, env...
  label=$1; shift
  env "$@" timeout -k 5 300 python bench.py --workload auni3d --no-e2e --no-cpu-baseline --no-extras --steps 10 > gpurun_out/r02w_x.json 2> gpurun_out/r02w_x.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02w_x.json') if l.startswith('{')][-1])
print('$label', 'auni3d ms', round(d['ms_per_step'],3), 'GCUPS', round(d['value'],3), 'launches/eval', d['gpu_launches']/(d['steps']*d['repeats']))
PY
}
{
run sequential AMPE_B200_KKS_PIPELINE=0
run pipe_default AMPE_B200_KKS_PIPELINE=1
run pipe_swap AMPE_B200_KKS_PIPELINE=1 AMPE_B200_PIPE_PRIO=swap
run pipe_kb64 AMPE_B200_KKS_PIPELINE=1 AMPE_B200_PIPE_KB=64
run pipe_kb256 AMPE_B200_KKS_PIPELINE=1 AMPE_B200_PIPE_KB=256
run pipe_ch64 AMPE_B200_KKS_PIPELINE=1 AMPE_B200_PIPE_CH=64
run pipe_swap_kb256 AMPE_B200_KKS_PIPELINE=1 AMPE_B200_PIPE_PRIO=swap AMPE_B200_PIPE_KB=256
run sequential AMPE_B200_KKS_PIPELINE=0
} 2>&1 | tee gpurun_out/r02w_ab2.log

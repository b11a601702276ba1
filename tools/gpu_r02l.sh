#!/bin/bash
# Round 2, call L (8 GPUs): weak scaling of the default workload (AuNi_3D 1024x1024x128 per GPU = 1024^3 on 8) and of
# Dendrite2D over NVSwitch, with the N = 1 lines of the same box; bit-exactness of the 8-rank ring on small grids.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
rm -f gpurun_out/r02l_bench.jsonl
timeout -k 5 300 python bench.py --no-e2e --no-cpu-baseline --no-extras >> gpurun_out/r02l_bench.jsonl 2> gpurun_out/r02l_n1.err
timeout -k 5 300 python bench.py --workload dendrite2d --no-e2e --no-cpu-baseline >> gpurun_out/r02l_bench.jsonl 2>> gpurun_out/r02l_n1.err
timeout -k 5 600 $TR bench.py --gpus $N --no-extras >> gpurun_out/r02l_bench.jsonl 2> gpurun_out/r02l_auni3d_n$N.err
timeout -k 5 300 $TR bench.py --gpus $N --workload dendrite2d --no-e2e >> gpurun_out/r02l_bench.jsonl 2> gpurun_out/r02l_dendrite2d_n$N.err
timeout -k 5 300 $TR bench.py --gpus $N --workload auni2d --no-e2e --no-extras >> gpurun_out/r02l_bench.jsonl 2> gpurun_out/r02l_auni2d_n$N.err
tail -3 gpurun_out/r02l_auni3d_n$N.err | cut -c1-300
python - <<PY
import json
for l in open('gpurun_out/r02l_bench.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); e=d.get('e2e')
        print(d['config']['workload'][:30], 'N', d['n_gpus'], 'ms %.4f GCUPS %.2f e2e %s launches/step %.1f transport %s'%(d['ms_per_step'], d['value'], e and round(e['value'],2), d['gpu_launches']/(d['steps']*d['repeats']), d['config'].get('halo_transport')))
PY
timeout -k 5 400 $TR tools/mgpu_check.py > gpurun_out/r02l_mgpu_check_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r02l_mgpu_check_n$N.log
grep -E "MGPU CHECK|MISMATCH|rc=" gpurun_out/r02l_mgpu_check_n$N.log | tail -5

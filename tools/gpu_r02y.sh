#!/bin/bash
# round 2, call Y (8 GPUs): final library -- weak scaling of the default workload and Dendrite2D with the N = 1 lines of
# the same box, 8-rank ring bit-exact (periodic and zero-slope cases), Dendrite deck on 8 slab ranks
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
rm -f gpurun_out/r02y_bench.jsonl
timeout -k 5 300 python bench.py --no-e2e --no-cpu-baseline --no-extras >> gpurun_out/r02y_bench.jsonl 2> gpurun_out/r02y_n1.err
timeout -k 5 300 python bench.py --workload dendrite2d --no-e2e --no-cpu-baseline >> gpurun_out/r02y_bench.jsonl 2>> gpurun_out/r02y_n1.err
timeout -k 5 600 $TR bench.py --gpus $N --no-extras >> gpurun_out/r02y_bench.jsonl 2> gpurun_out/r02y_auni3d_n$N.err
timeout -k 5 300 $TR bench.py --gpus $N --workload dendrite2d --no-e2e >> gpurun_out/r02y_bench.jsonl 2> gpurun_out/r02y_dendrite2d_n$N.err
tail -3 gpurun_out/r02y_auni3d_n$N.err | cut -c1-300
python - <<PY
import json
for l in open('gpurun_out/r02y_bench.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); e=d.get('e2e')
        print(d['config']['workload'][:30], 'N', d['n_gpus'], 'ms %.4f GCUPS %.2f e2e %s launches/step %.1f transport %s cpus/rank %s'%(d['ms_per_step'], d['value'], e and round(e['value'],2), d['gpu_launches']/(d['steps']*d['repeats']), d['config'].get('halo_transport'), d['config'].get('host_cpus_per_rank')))
PY
MGPU_MODES=default timeout -k 5 400 $TR tools/mgpu_check.py > gpurun_out/r02y_mgpu_check_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r02y_mgpu_check_n$N.log
grep -E "MGPU CHECK|MISMATCH|rc=" gpurun_out/r02y_mgpu_check_n$N.log | tail -5
timeout -k 5 300 $TR tools/mgpu_deck.py 0 > gpurun_out/r02y_mgpu_deck_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r02y_mgpu_deck_n$N.log
grep -E "^rank 0|MGPU DECK|rc=" gpurun_out/r02y_mgpu_deck_n$N.log | tail -4
lscpu | grep -i "numa\|socket\|^CPU(s)" > gpurun_out/r02y_topo.txt; free -g | head -2 >> gpurun_out/r02y_topo.txt; cat gpurun_out/r02y_topo.txt

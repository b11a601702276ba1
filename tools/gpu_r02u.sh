#!/bin/bash
# round 2, call U (2 GPUs): slab == single GPU with the zero-slope cases over real peer links, the regression decks on two
# slab ranks (one GPU each), the default bench at N = 2 with e2e (ranks bound next to their GPUs)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout -k 5 600 $TR tools/mgpu_check.py > gpurun_out/r02u_mgpu_check_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r02u_mgpu_check_n$N.log
grep -v "^W\|^\[W\|^\*\*\*\|OMP_NUM" gpurun_out/r02u_mgpu_check_n$N.log | tail -12
timeout -k 5 900 $TR tools/mgpu_deck.py > gpurun_out/r02u_mgpu_deck_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r02u_mgpu_deck_n$N.log
grep -E "^rank 0|MGPU|rc=" gpurun_out/r02u_mgpu_deck_n$N.log
timeout -k 5 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02u_bench_n$N.json 2> gpurun_out/r02u_bench_n$N.err
python - <<PY
import json
for l in open('gpurun_out/r02u_bench_n$N.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['config']['workload'][:30], 'N', d['n_gpus'], 'ms %.4f GCUPS %.2f e2e %.3f host cpus/rank %s'%(d['ms_per_step'], d['value'], d['e2e']['value'], d['config'].get('host_cpus_per_rank')))
PY
nvidia-smi topo -m > gpurun_out/r02u_topo.txt 2>&1; lscpu | grep -i "numa\|socket\|^CPU(s)" >> gpurun_out/r02u_topo.txt
tail -12 gpurun_out/r02u_topo.txt

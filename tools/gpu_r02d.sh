#!/bin/bash
# Round 2, call D (N GPUs, default 2): bit-exactness of the slab evaluation over real peer links, then bench lines
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout -k 5 600 $TR tools/mgpu_check.py > gpurun_out/r02d_mgpu_check_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r02d_mgpu_check_n$N.log
grep -v "^W\|^\[W" gpurun_out/r02d_mgpu_check_n$N.log | tail -14
g++ -std=c++17 -O1 -I include -I /usr/local/cuda/include tests/cpp/halo_two_ranks.cpp -L ampe_b200 -lampe_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/ampe_b200 -o /tmp/halo_two_ranks && /tmp/halo_two_ranks 2 | tail -3
for w in auni3d dendrite2d auni2d gg3d_hbsm; do
  timeout -k 5 400 $TR bench.py --gpus $N --workload $w --steps 20 --warmup 5 >> gpurun_out/r02d_bench_n$N.jsonl 2> gpurun_out/r02d_bench_${w}_n$N.err
  tail -2 gpurun_out/r02d_bench_${w}_n$N.err | cut -c1-300
done
python - <<PY
import json
for l in open('gpurun_out/r02d_bench_n$N.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print(d['config']['workload'][:30], 'N', d['n_gpus'], 'ms %.4f GCUPS %.2f e2e %.2f launches/step %d transport %s'%(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches']//(d['steps']*d['repeats']), d['config'].get('halo_transport')))
PY

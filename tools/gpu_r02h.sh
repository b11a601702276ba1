#!/bin/bash
# Round 2, call H (2 GPUs): the in-kernel ghost-plane wait over real peer links (bit-exactness), then Dendrite2D N = 1 | 2
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout -k 5 600 $TR tools/mgpu_check.py > gpurun_out/r02h_mgpu_check_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r02h_mgpu_check_n$N.log
grep -v "^W\|^\[W\|^\*\*\*\|OMP_NUM" gpurun_out/r02h_mgpu_check_n$N.log | tail -18
for w in dendrite2d pfhub1a; do
  timeout -k 5 300 python bench.py --workload $w --steps 20 --warmup 5 --no-e2e --no-cpu-baseline >> gpurun_out/r02h_bench.jsonl 2> gpurun_out/r02h_bench_${w}_n1.err
  timeout -k 5 400 $TR bench.py --gpus $N --workload $w --steps 20 --warmup 5 --no-e2e >> gpurun_out/r02h_bench.jsonl 2> gpurun_out/r02h_bench_${w}_n$N.err
  AMPE_B200_HALO_INKERNEL=0 timeout -k 5 400 $TR bench.py --gpus $N --workload $w --steps 20 --warmup 5 --no-e2e >> gpurun_out/r02h_bench.jsonl 2>> gpurun_out/r02h_bench_${w}_n$N.err
  tail -2 gpurun_out/r02h_bench_${w}_n$N.err | cut -c1-300
done
python - <<PY
import json
for l in open('gpurun_out/r02h_bench.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print(d['config']['workload'][:30], 'N', d['n_gpus'], 'ms %.4f GCUPS %.2f launches/step %.1f'%(d['ms_per_step'], d['value'], d['gpu_launches']/(d['steps']*d['repeats'])))
PY

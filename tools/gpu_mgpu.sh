#!/bin/bash
# N-GPU round: NCCL slab check + bench lines at N ranks (usage: gpu_mgpu.sh N).  Every step is
# guarded by a short timeout: a hung collective must not eat the GPU budget.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout -k 5 ${T_CHECK:-120} $TR tools/mgpu_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -12
echo "check rc=${PIPESTATUS[0]}"
for w in ${WORKLOADS:-dendrite2d auni3d}; do
  timeout -k 5 ${T_BENCH:-150} $TR bench.py --gpus $N --workload $w --steps 20 --warmup 3 > gpurun_out/bench_${w}_n$N.json 2> gpurun_out/bench_${w}_n$N.err
  echo "bench $w rc=$?"; tail -c 1500 gpurun_out/bench_${w}_n$N.json; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_${w}_n$N.err | tail -3
done
if [ -n "$TRY_GRAPHS" ]; then
  echo "== CUDA-graph replay of the exchange"
  AMPE_B200_GRAPHS=1 timeout -k 5 60 $TR tools/mgpu_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -8
  echo "graph check rc=${PIPESTATUS[0]}"; tail -2 gpurun_out/mgpu_progress_rank0.log
fi

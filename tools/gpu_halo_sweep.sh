#!/bin/bash
# exchange variants at N ranks on the latency-bound 2D workload (usage: gpu_halo_sweep.sh N)
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for order in interior exchange; do for mode in p2p allgather; do for prio in 1 0; do
  echo "== order=$order mode=$mode nccl_high_priority=$prio"
  AMPE_B200_HALO_ORDER=$order AMPE_B200_HALO_MODE=$mode TORCH_NCCL_HIGH_PRIORITY=$prio timeout -k 5 60 $TR bench.py --gpus $N --steps 100 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   ms/step %.4f  GCUPS %.2f'%(d['ms_per_step'], d['value']))"
done; done; done

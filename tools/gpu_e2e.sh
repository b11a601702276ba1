#!/bin/bash
# pipelined host entry point: parity tests + e2e bench lines at several chunk counts
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for ch in 1 4 8 16; do
  for w in dendrite2d auni3d; do
    echo "== chunks $ch $w"
    AMPE_B200_HOST_CHUNKS=$ch timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f e2e %s'%(d['value'], d['e2e']))"
  done
done

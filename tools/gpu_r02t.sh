#!/bin/bash
# round 2, call T: zero-slope boundaries along the slab axis on N ranks, implicit integrator on slab ranks (regression decks)
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests/test_gpu_halo_two_ranks.py -q -m gpu -x --durations=5 -k regression_decks > gpurun_out/r02t_pytest_slab.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02t_pytest_slab.log
tail -30 gpurun_out/r02t_pytest_slab.log | grep -E "rank [0-9]|MGPU|passed|failed"
timeout -k 5 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29655 tools/mgpu_deck.py > gpurun_out/r02t_mgpu_deck.log 2>&1
grep -E "^rank|MGPU" gpurun_out/r02t_mgpu_deck.log | tail -20

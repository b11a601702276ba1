#!/bin/bash
# Round 2, call A: everything that had never run on a GPU (default suite + the opt-in tests), V-cycle timing
# and ncu of the multigrid, the split-3D A/B, baseline bench lines of all workloads, ncu captures of C3 / C4.
mkdir -p gpurun_out
bash tools/gpu_round2_first.sh s34
for w in dendrite2d auni2d gg3d_hbsm auni3d pfhub1a; do
  timeout -k 5 200 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline >> gpurun_out/r02a_bench_lines.jsonl 2> gpurun_out/bench_$w.err
done
cat gpurun_out/r02a_bench_lines.jsonl | cut -c1-400
bash tools/gpu_prof.sh r02a auni2d gg3d_hbsm
ls -la gpurun_out | tail -12

#!/bin/bash
# ncu --set full capture of the fused kernel for the given workloads -> gpurun_out/prof_<tag>_<w>.ncu-rep
tag=$1; shift
mkdir -p gpurun_out
for w in "$@"; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rhs_ -s 3 -c 1 -f -o gpurun_out/prof_${tag}_$w \
  python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_${tag}_$w.log 2>&1
done

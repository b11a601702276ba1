#!/bin/bash
# round 2, call P: 3D anisotropic phase flux in the fused marching kernel -- parity cases, slab split, bench unchanged
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_strategies.py -q -m gpu -k "model_switches or slab_decomposition or anisotropic or error_paths" --durations=5 2>&1 | tail -25 | tee gpurun_out/r02p_pytest_aniso3d.log
echo "pytest rc=${PIPESTATUS[0]}"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02p_bench_auni3d.json 2> gpurun_out/r02p_bench_auni3d.err
tail -c 1500 gpurun_out/r02p_bench_auni3d.json

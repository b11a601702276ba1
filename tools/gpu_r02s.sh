#!/bin/bash
# round 2, call S: compute-sanitizer over the small parity cases (memcheck, racecheck, synccheck, initcheck)
mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
SEL='rhs_matches_oracle or slab_decomposition or zero_slope or anisotropic_3d or fd_flag_lagging'
for tool in memcheck racecheck synccheck; do
  timeout -k 5 1500 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "$SEL" > gpurun_out/r02s_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/r02s_summary.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r02s_$tool.log | tail -3 | tee -a gpurun_out/r02s_summary.log
done
# multigrid + vector kernels + energy diagnostics under memcheck and racecheck
for tool in memcheck racecheck; do
  timeout -k 5 1500 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 20 \
    python -m pytest tests/test_gpu_round2_regressions.py tests/test_gpu_widening.py -q -m gpu -x > gpurun_out/r02s_${tool}_widening.log 2>&1
  echo "$tool widening rc=$?" | tee -a gpurun_out/r02s_summary.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r02s_${tool}_widening.log | tail -3 | tee -a gpurun_out/r02s_summary.log
done

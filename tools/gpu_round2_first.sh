#!/bin/bash
# FIRST gpurun call of round 2 (everything written after round 1's GPU budget ran out, in one call):
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh'
# 1. the default GPU suite (must stay green), 2. the opt-in tests of the code that has never run on a GPU
# (implicit-integrator failure code, split 3D launches, block preconditioners), 3. V-cycle timing + ncu of the
# multigrid kernels, 4. the split-3D A/B.  Every step has its own timeout; logs land in gpurun_out/.
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
AMPE_B200_RUN_EXPERIMENTS=1 timeout -k 5 600 python -m pytest tests -m gpu -q \
  -k "precond or split3d or newton_failure" > gpurun_out/pytest_experiments.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_experiments.log
tail -25 gpurun_out/pytest_experiments.log
bash tools/gpu_precond.sh
bash tools/gpu_split3d.sh "$@" > gpurun_out/split3d_ab.log 2>&1; tail -12 gpurun_out/split3d_ab.log

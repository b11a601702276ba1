#!/bin/bash
# The reference's own regression test scripts (tests/<deck>/test2d.py), UNMODIFIED, with the device program as the executable:
#   test2d.py <mpiexec> <-n> <np> <exe> <input> [<thermo data dir>]     (the CTest command line of the reference)
# _ref_scratch/ holds the scripts, decks and generators copied from /root/reference for this one call (git-ignored, deleted
# afterwards: it only exists because /root/reference does not exist on the GPU box).
mkdir -p gpurun_out
LOG=$PWD/gpurun_out/${LOGNAME_R02:-r02ak_reference_test_scripts.log}
: > $LOG
export PYTHONPATH=$PWD:$PWD/tools/netcdf4_shim:$PYTHONPATH
ROOT=$PWD
for case in "$@"; do
  deck=${case%%:*}; script=${case##*:}; dim=${script#test}; dim=${dim%.py}
  cd $ROOT/_ref_scratch/tests/$deck
  echo "=== tests/$deck/$script  exe = python -m ampe_b200.run_deck" >> $LOG
  start=$(date +%s.%N)
  timeout ${DECK_TIMEOUT:-120} python $script "" "" "" "python -m ampe_b200.run_deck" $dim.input $ROOT/_ref_scratch/thermo >> $LOG 2>&1
  rc=$?
  echo "=== tests/$deck/$script exit code $rc  ($(python -c "import time; print('%.1f s' % (time.time() - $start))"))" >> $LOG
  echo "tests/$deck/$script exit code $rc"
done

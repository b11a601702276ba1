#!/bin/bash
# build variants/lib_<name>.so with extra nvcc flags (tile-shape sweeps etc.; development aid)
# usage: tools/build_variant.sh name "-DAMPE_T3Y=8 ..."
set -e
name=$1; shift
flags="$*"
root=$(cd $(dirname $0)/.. && pwd)
out=$root/variants/$name
mkdir -p $out
cd $root/ampe_b200/csrc
pids=""
for f in $(ls *.cu | sed 's/\.cu$//'); do
  if [ -f $f.cu ]; then
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=${FMAD:-false} -std=c++17 --extended-lambda -Xcompiler -fPIC $flags -c -o $out/$f.o $f.cu &
  pids="$pids $!"
  fi
done
for p in $pids; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $root/variants/lib_$name.so $out/*.o
rm -rf $out
echo built variants/lib_$name.so

#!/bin/bash
# Build kernel variants with different tuning knobs into build/ (run on the CPU box), e.g.
#   tools/sweep.sh "REFILL=8 LEAF_BATCH=8 STEPS=4 MIN_BLOCKS=8" "REFILL=16 ..."
set -e
cd "$(dirname "$0")/../libyafaray_b200/csrc"
mkdir -p ../../build
for v in "$@"; do
  defs=""; name="v"
  for kv in $v; do defs="$defs -DB200RT_$kv"; name="${name}_${kv//=/}"; done
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O3 $defs -shared -o ../../build/$name.so b200rt.cu kd_build.cc -lpthread &
done
wait
ls ../../build

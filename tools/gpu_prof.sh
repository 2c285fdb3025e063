#!/bin/bash
# ncu --set full capture of the closest + shadow kernels of tools/quick_bench.py (one launch each, after warm-up)
#   gpurun -- 'bash tools/gpu_prof.sh <tag> [lib.so] [quick_bench args]'
tag=${1:-prof}; lib=${2:-}; shift; shift
mkdir -p gpurun_out
B200RT_LIB=$lib timeout 900 ncu --set full --clock-control none --import-source on -k regex:traceKernel -s 6 -c 2 -f -o gpurun_out/${tag}_prof \
    python tools/quick_bench.py --steps 1 "$@" > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log

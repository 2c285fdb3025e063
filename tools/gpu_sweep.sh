#!/bin/bash
# Kernel-knob sweep on the GPU box: every build/v_*.so produced by tools/sweep.sh through tools/quick_bench.py (S1M-hf, 16 Mi rays).
#   tools/sweep.sh "LEAF_BATCH=12" ... && gpurun --timeout 900 -- 'bash tools/gpu_sweep.sh <tag>'
tag=${1:-sweep}
mkdir -p gpurun_out
timeout 120 python tools/quick_bench.py --tag default > gpurun_out/${tag}_knobs.txt 2>&1
for lib in build/v_*.so; do
  B200RT_LIB=$lib timeout 120 python tools/quick_bench.py --tag $(basename $lib .so) 2>&1 | tail -1 >> gpurun_out/${tag}_knobs.txt
done
timeout 120 python tools/quick_bench.py --tag default-again 2>&1 | tail -1 >> gpurun_out/${tag}_knobs.txt
cat gpurun_out/${tag}_knobs.txt

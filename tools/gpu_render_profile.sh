#!/bin/bash
# Sampling profiles (render_bench B200_PROF=1) of a path-traced and a direct-lighting frame, stock kd-tree vs b200-kdtree.
#   gpurun --timeout 600 -- 'bash tools/gpu_render_profile.sh <tag>'
tag=${1:-rprof}
mkdir -p gpurun_out; cd gpurun_out
for integ in pathtracing directlighting; do
  aa=8; [ $integ = directlighting ] && aa=4
  for accel in yafaray-kdtree-original b200-kdtree; do
    B200_PROF=1 timeout 200 ../integration/_build/render_bench $accel $integ 707 1920 1080 $aa /tmp/${tag}.tga -1 wavefront_fibers=512 wavefront_block=2 wavefront_groups=2 > ${tag}_${integ}_${accel}.log 2>&1
    grep -h "RENDER_BENCH\|wavefront rays" ${tag}_${integ}_${accel}.log | cut -c1-330
  done
done

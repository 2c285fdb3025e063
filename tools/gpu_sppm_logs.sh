#!/bin/bash
# SPPM on the GPU box with the reference's own log lines and render_bench's sampling profile (B200_PROF): where does a
# b200-kdtree SPPM frame spend its CPU time?
#   gpurun --timeout 600 -- 'bash tools/gpu_sppm_logs.sh <tag>'
tag=${1:-sppm}
mkdir -p gpurun_out
cd gpurun_out
run() { # label accel extra...
  label=$1; accel=$2; shift 2
  for k in 1 2; do
    B200_PROF=1 timeout 200 ../integration/_build/render_bench $accel SPPM 707 960 540 1 /tmp/${tag}_$label.tga -1 i:photons=500000 i:passNums=2 "$@" > ${tag}_${label}_run$k.log 2>&1
    grep -h "RENDER_BENCH" ${tag}_${label}_run$k.log | sed "s/^RENDER_BENCH {/{\"run\": \"$label\", /" | cut -c1-330
  done
}
run stock yafaray-kdtree-original
run b200 b200-kdtree wavefront_fibers=512 wavefront_block=2 wavefront_groups=2



B200_MIN_PHOTONS_PER_WORKER=256 run b200_mp256 b200-kdtree wavefront_fibers=512 wavefront_block=2 wavefront_groups=2
grep -h "PhotonMap building time" ${tag}_*_run2.log

#!/bin/bash
# Other regimes and the photon renders on one GPU: bench.py on the HBM-resident 10 M-triangle scene and on coherent camera rays, an
# ncu capture of the 10 M-triangle traversal, SPPM / photon mapping renders against the stock kd-tree.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round_regimes.sh <tag>'
tag=${1:-regimes}
mkdir -p gpurun_out
timeout 400 python bench.py --workload rcoh --steps 5 --warmup 3 --cpu-seconds 4 > gpurun_out/${tag}_bench_rcoh.json 2> gpurun_out/${tag}_bench.err; tail -c 400 gpurun_out/${tag}_bench_rcoh.json
timeout 600 python bench.py --workload s10m --steps 5 --warmup 3 --no-cpu-baseline --no-tshadow > gpurun_out/${tag}_bench_s10m.json 2>> gpurun_out/${tag}_bench.err; tail -c 400 gpurun_out/${tag}_bench_s10m.json
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:traceKernel|setupKernel" -s 16 -c 4 -f -o gpurun_out/${tag}_s10m_prof \
    python bench.py --workload s10m --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-tshadow > gpurun_out/${tag}_s10m_ncu.log 2>&1
for integ in SPPM photonmapping; do
  extra="i:diffuse_photons=1000000 i:caustic_photons=200000 b:finalGather=0"; [ $integ = SPPM ] && extra="i:photons=500000 i:passNums=2"
  for rep in 1 2; do
    timeout 300 python tools/render_compare.py --integrator $integ --width 960 --height 540 --aa 1 --fibers 512 --block 2 --skip-second-stock --extra "$extra" 2>> gpurun_out/${tag}_bench.err >> gpurun_out/${tag}_render_photon.jsonl
  done
done
cut -c1-330 gpurun_out/${tag}_render_photon.jsonl

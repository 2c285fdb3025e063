#!/bin/bash
# Other regimes and the render path on one GPU: the whole -m gpu suite (all failures, not just the first), bench.py on the
# HBM-resident 10 M-triangle scene and on coherent camera rays, an ncu capture of the 10 M-triangle traversal, SPPM / photon
# mapping renders with different photon-worker granularities.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round_regimes.sh <tag>'
tag=${1:-regimes}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -6 gpurun_out/${tag}_pytest.log
cp gpurun_out/parity_report.jsonl gpurun_out/${tag}_parity_report.jsonl 2>/dev/null
timeout 400 python bench.py --workload rcoh --steps 5 --warmup 3 --cpu-seconds 4 > gpurun_out/${tag}_bench_rcoh.json 2> gpurun_out/${tag}_bench.err; tail -c 600 gpurun_out/${tag}_bench_rcoh.json
timeout 600 python bench.py --workload s10m --steps 5 --warmup 3 --no-cpu-baseline --no-tshadow > gpurun_out/${tag}_bench_s10m.json 2>> gpurun_out/${tag}_bench.err; tail -c 600 gpurun_out/${tag}_bench_s10m.json
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:traceKernel|setupKernel" -s 12 -c 4 -f -o gpurun_out/${tag}_s10m_prof \
    python bench.py --workload s10m --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-tshadow > gpurun_out/${tag}_s10m_ncu.log 2>&1
for integ in SPPM photonmapping; do
  for mp in 16 64 256 1024; do
    extra="i:diffuse_photons=1000000 i:caustic_photons=200000 b:finalGather=0"; [ $integ = SPPM ] && extra="i:photons=500000 i:passNums=2"
    B200_MIN_PHOTONS_PER_WORKER=$mp timeout 300 python tools/render_compare.py --integrator $integ --width 960 --height 540 --aa 1 --fibers 512 --block 2 --skip-second-stock --extra "$extra" 2>> gpurun_out/${tag}_bench.err | sed "s/^{/{\"min_photons_per_worker\": $mp, /" >> gpurun_out/${tag}_render_photon.jsonl
  done
done
cut -c1-400 gpurun_out/${tag}_render_photon.jsonl

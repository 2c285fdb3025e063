#!/bin/bash
# Final visit of a round: everything gpu_round.sh does, then the PCIe ceiling of the e2e arm and the photon-integrator renders.
tag=${1:-final}
bash tools/gpu_round.sh $tag
timeout 120 python tools/pcie_peak.py > gpurun_out/${tag}_pcie.json 2>&1; cat gpurun_out/${tag}_pcie.json
for integ in photonmapping SPPM; do
  extra="i:diffuse_photons=1000000 i:caustic_photons=200000 b:finalGather=0"; [ $integ = SPPM ] && extra="i:photons=500000 i:passNums=2"
  timeout 600 python tools/render_compare.py --integrator $integ --width 960 --height 540 --aa 1 --fibers 512 --block 2 --groups 2 --skip-second-stock --extra "$extra" >> gpurun_out/${tag}_render_photon.jsonl 2>&1
done
cut -c1-700 gpurun_out/${tag}_render_photon.jsonl

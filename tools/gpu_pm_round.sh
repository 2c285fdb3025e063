#!/bin/bash
# Photon-map gather on the GPU box: parity tests, throughput (heaps in shared memory vs in the result array), one ncu capture.
#   gpurun -- 'bash tools/gpu_pm_round.sh <tag>'
tag=${1:-pm}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_pm.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 100 python tools/pm_bench.py > gpurun_out/${tag}_pm_bench.json 2> gpurun_out/${tag}_pm_bench.err; cut -c1-1500 gpurun_out/${tag}_pm_bench.json; tail -2 gpurun_out/${tag}_pm_bench.err
B200PM_SMEM_K=0 timeout 60 python tools/pm_bench.py --no-cpu > gpurun_out/${tag}_pm_bench_globalheap.json 2>> gpurun_out/${tag}_pm_bench.err; cut -c1-400 gpurun_out/${tag}_pm_bench_globalheap.json
timeout 60 python tools/pm_bench.py --no-cpu --k 8 --sq-radius 1e-4 > gpurun_out/${tag}_pm_bench_k8.json 2>> gpurun_out/${tag}_pm_bench.err; cut -c1-400 gpurun_out/${tag}_pm_bench_k8.json
timeout 120 ncu --set full --clock-control none --import-source on -k regex:pmLookupKernel -s 3 -c 1 -f -o gpurun_out/${tag}_pm_prof \
    python tools/pm_bench.py --no-cpu --steps 1 > gpurun_out/${tag}_pm_ncu.log 2>&1
tail -2 gpurun_out/${tag}_pm_ncu.log | cut -c1-300

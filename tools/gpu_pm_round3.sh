#!/bin/bash
# Third photon-map visit: every kernel variant against the oracle, the variant sweep, one ncu capture of the phased + single-pop kernel.
tag=${1:-pm3}
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_pm.py -m gpu -x -q -k "variants or golden" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 150 python tools/pm_sweep.py > gpurun_out/${tag}_pm_sweep.jsonl 2> gpurun_out/${tag}_pm_sweep.err; cut -c40-240 gpurun_out/${tag}_pm_sweep.jsonl; tail -3 gpurun_out/${tag}_pm_sweep.err
B200PM_KERNEL=phased timeout 100 ncu --set full --clock-control none --import-source on -k regex:pmLookup -s 3 -c 1 -f -o gpurun_out/${tag}_pm_prof \
    python tools/pm_bench.py --no-cpu --steps 1 > gpurun_out/${tag}_pm_ncu.log 2>&1
tail -2 gpurun_out/${tag}_pm_ncu.log | cut -c1-300

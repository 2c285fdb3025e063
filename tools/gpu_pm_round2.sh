#!/bin/bash
# Second photon-map visit: parity tests with the phased kernel as the default, the variant sweep, pm_bench (pinned e2e), one ncu capture.
tag=${1:-pm2}
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_pm.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 120 python tools/pm_sweep.py > gpurun_out/${tag}_pm_sweep.jsonl 2> gpurun_out/${tag}_pm_sweep.err; cut -c1-230 gpurun_out/${tag}_pm_sweep.jsonl; tail -3 gpurun_out/${tag}_pm_sweep.err
timeout 100 python tools/pm_bench.py > gpurun_out/${tag}_pm_bench.json 2> gpurun_out/${tag}_pm_bench.err; cut -c1-1800 gpurun_out/${tag}_pm_bench.json; tail -2 gpurun_out/${tag}_pm_bench.err
timeout 100 ncu --set full --clock-control none --import-source on -k regex:pmLookup -s 3 -c 1 -f -o gpurun_out/${tag}_pm_prof \
    python tools/pm_bench.py --no-cpu --steps 1 > gpurun_out/${tag}_pm_ncu.log 2>&1
tail -2 gpurun_out/${tag}_pm_ncu.log | cut -c1-300

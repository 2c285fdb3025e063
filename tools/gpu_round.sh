#!/bin/bash
# One GPU-box visit: smoke(), parity tests, both bench arms, the ncu launch list of the bench command and one full
# ncu capture of the closest + shadow kernels.  Outputs under gpurun_out/ (summaries are copied to profiles/).
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'
tag=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
cp gpurun_out/parity_report.jsonl gpurun_out/${tag}_parity_report.jsonl 2>/dev/null
tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 3000 gpurun_out/${tag}_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err; tail -c 1500 gpurun_out/${tag}_bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-tshadow > gpurun_out/${tag}_ncu_list.log 2>&1
# warm-up = 3 steps x (setup + traversal) x (closest + shadow) = 12 launches; then one of each
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:traceKernel|setupKernel" -s 12 -c 4 -f -o gpurun_out/${tag}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-tshadow > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out

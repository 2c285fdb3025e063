#!/bin/bash
# Last visit of the round: smoke(), the whole GPU suite, the judged bench line (with the photon_gather key).  No profiler.
tag=${1:-last}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log | cut -c1-200
rm -f gpurun_out/parity_report.jsonl
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
cp gpurun_out/parity_report.jsonl gpurun_out/${tag}_parity_report.jsonl 2>/dev/null
tail -4 gpurun_out/${tag}_pytest.log
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -2 gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["roofline"]["bound"], d["roofline"]["frac"])
print(json.dumps(d.get("photon_gather"))[:1500])
PY

#!/bin/bash
# e2e (host-buffer C-ABI calls) against the staging chunk size and the number of chunks in flight.
tag=${1:-e2e}
mkdir -p gpurun_out
for cfg in "32 3" "8 3" "16 3" "64 3" "16 4" "8 4" "4 4" "16 6"; do set -- $cfg
  B200RT_CHUNK_MB=$1 B200RT_LANES=$2 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
r = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk_mb $1 lanes $2 e2e', round(r['e2e']['value'], 1), 'device', round(r['value'], 1))" >> gpurun_out/${tag}_e2e_sweep.txt
done
cat gpurun_out/${tag}_e2e_sweep.txt

#!/bin/bash
# compute-sanitizer over the parity tests that reach every kernel variant (golden vectors incl. quads / spheres / motion blur, job
# bundles = the mixed-kind kernel, one-ray calls, deep transparent shadows, a two-pass batch): memcheck, then racecheck and synccheck
# (shared-memory hazards and barrier use of the TMA-fed setup pass).
#   gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh <tag>'
tag=${1:-sanitize}
mkdir -p gpurun_out
sel="golden_vectors or trace_jobs_bundles or single_ray_calls or empty_scene or deeper_than or tree_space_rays or host_and_device"
{ timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu.py -m gpu -q -k "$sel" 2>&1 | tail -6; echo "memcheck exit $?"; } > gpurun_out/${tag}_memcheck.txt
tail -4 gpurun_out/${tag}_memcheck.txt
{ timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu.py -m gpu -q -k "golden_vectors or host_and_device" 2>&1 | tail -6; echo "racecheck exit $?"; } > gpurun_out/${tag}_racecheck.txt
tail -4 gpurun_out/${tag}_racecheck.txt
{ timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu.py -m gpu -q -k "golden_vectors or host_and_device" 2>&1 | tail -6; echo "synccheck exit $?"; } > gpurun_out/${tag}_synccheck.txt
tail -4 gpurun_out/${tag}_synccheck.txt

#!/bin/bash
# The flush combiner on the GPU box: the job / render tests, then path-traced and direct-lighting frames with and without it and
# with smaller fiber groups (what the merged launches are for).
#   gpurun --timeout 1200 -- 'bash tools/gpu_combiner.sh <tag>'
tag=${1:-comb}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu.py tests/test_film.py tests/test_render.py -m gpu -q -x -k "jobs or concurrent or single_ray or render or photon or shards or instances or motion or sphere or material" > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
run() { # label env... -- args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python tools/render_compare.py --width 1920 --height 1080 --block 2 --skip-second-stock "$@" 2>>gpurun_out/${tag}.err | grep b200 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); w = d.get('wavefront', {})
    print(json.dumps({'label': '$label', 'integrator': d['integrator'], 'run': d['run'], 'render_seconds': d['render_seconds'], 'rays_per_batch': w.get('rays_per_batch'), 'rays_per_launch': w.get('rays_per_launch'), 'kernel_launches': w.get('kernel_launches'), 'trace_thread_seconds': w.get('trace_thread_seconds'), 'worker_thread_seconds': w.get('worker_thread_seconds'), 'psnr': d.get('psnr_vs_first_stock_db')}))
" | tee -a gpurun_out/${tag}_renders.jsonl
}
timeout 200 python tools/render_compare.py --width 1920 --height 1080 --integrator pathtracing --aa 8 --fibers 512 --skip-second-stock 2>/dev/null | grep '"stock"' | cut -c1-260 | tee -a gpurun_out/${tag}_renders.jsonl
run off B200RT_COMBINE=0 -- --integrator pathtracing --aa 8 --fibers 512,256 --groups 2
run r2048 B200RT_COMBINE=1 -- --integrator pathtracing --aa 8 --fibers 512,256,128 --groups 2
run r2048g4 B200RT_COMBINE=1 -- --integrator pathtracing --aa 8 --fibers 256,512 --groups 4
run r1024 B200RT_COMBINE_RAYS=1024 -- --integrator pathtracing --aa 8 --fibers 512,256,128 --groups 2
run r4096 B200RT_COMBINE_RAYS=4096 -- --integrator pathtracing --aa 8 --fibers 512,256 --groups 2
timeout 200 python tools/render_compare.py --width 1920 --height 1080 --integrator directlighting --aa 4 --fibers 512 --skip-second-stock 2>/dev/null | grep '"stock"' | cut -c1-260 | tee -a gpurun_out/${tag}_renders.jsonl
run off B200RT_COMBINE=0 -- --integrator directlighting --aa 4 --fibers 512,256 --groups 2
run r2048 B200RT_COMBINE=1 -- --integrator directlighting --aa 4 --fibers 512,256,128 --groups 2
run r1024 B200RT_COMBINE_RAYS=1024 -- --integrator directlighting --aa 4 --fibers 512,256,128 --groups 2

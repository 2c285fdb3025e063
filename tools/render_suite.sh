#!/bin/bash
# End-to-end yafaray_render comparisons (stock CPU kd-tree vs b200-kdtree wavefront) on the GPU box; see tools/render_compare.py.
out=gpurun_out/render_suite.jsonl
rm -f $out
run() { timeout 1200 python tools/render_compare.py --skip-second-stock --fibers 512 --groups 2 "$@" >> $out 2>&1; }
run --cells 707 --width 960 --height 540 --aa 4 --integrator directlighting --extra "b:do_AO=1 i:AO_samples=32 f:AO_distance=3"
run --cells 2236 --width 960 --height 540 --aa 4 --integrator directlighting --extra "b:do_AO=1 i:AO_samples=32 f:AO_distance=3"
python - <<PY
import json
for l in open("$out"):
    try: r = json.loads(l)
    except Exception: print(l[:300]); continue
    w = r.get("wavefront", {})
    print(r.get("run"), r.get("integrator"), r.get("triangles"), "pre_s", r.get("preprocess_seconds"), "render_s", r.get("render_seconds"), "psnr", round(r.get("psnr_vs_first_stock_db", 0), 1),
          "rays", w.get("closest", 0) + w.get("shadow", 0), "rays/batch", w.get("rays_per_batch"), "trace_ts", w.get("trace_thread_seconds"), "worker_ts", w.get("worker_thread_seconds"))
PY

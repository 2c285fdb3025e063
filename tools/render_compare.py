#!/usr/bin/env python
"""End-to-end `yafaray_render` comparison on the GPU box (BASELINE.json configs[2]-like scene):
integration/_build/render_bench (a client of libYafaRay's public C API) renders the synthetic height-field scene with
the stock CPU kd-tree and with the b200-kdtree accelerator (wavefront ray queue), same integrator, same threads; prints
one JSON line per run plus the PSNR of every image against the first stock run (a second stock run gives the noise
floor of the comparison: the reference seeds its tile RNGs from rand() and thread timing).

    python tools/render_compare.py [--cells 707] [--width 960 --height 540] [--aa 4] [--integrator pathtracing] [--threads -1]
"""
import argparse, json, os, re, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_render import read_tga, psnr

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=707)
ap.add_argument("--width", type=int, default=960)
ap.add_argument("--height", type=int, default=540)
ap.add_argument("--aa", type=int, default=4)
ap.add_argument("--integrator", default="pathtracing")
ap.add_argument("--threads", type=int, default=-1)
ap.add_argument("--fibers", default="1024")
ap.add_argument("--block", default="4")
ap.add_argument("--groups", default="2")
ap.add_argument("--skip-second-stock", action="store_true")
ap.add_argument("--extra", default="", help="extra render_bench arguments for every run, e.g. 'b:do_AO=1 i:AO_samples=32'")
ap.add_argument("--per-ray", action="store_true", help="also time the per-ray compatibility path (wavefront_fibers=0); slow")
a = ap.parse_args()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
binary = os.path.join(root, "integration", "_build", "render_bench")
runs = [("stock", "yafaray-kdtree-original", [])]
if not a.skip_second_stock:
    runs.append(("stock-again", "yafaray-kdtree-original", []))
for f in a.fibers.split(","):
    for b in a.block.split(","):
        for g in a.groups.split(","):
            runs.append((f"b200-f{f}-b{b}-g{g}", "b200-kdtree", [f"wavefront_fibers={f}", f"wavefront_block={b}", f"wavefront_groups={g}"]))
if a.per_ray:
    runs.append(("b200-per-ray", "b200-kdtree", ["wavefront_fibers=0"]))
first = None
with tempfile.TemporaryDirectory() as d:
    for name, accel, extra in runs:
        out = os.path.join(d, name + ".tga")
        cmd = [binary, accel, a.integrator, str(a.cells), str(a.width), str(a.height), str(a.aa), out, str(a.threads)] + extra + a.extra.split()
        p = subprocess.run(cmd, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, errors="replace")
        m = re.search(r"RENDER_BENCH (\{.*\})", p.stdout)
        rec = json.loads(m.group(1)) if m else {"error": p.stdout[-2000:]}
        rec["run"] = name
        wf = re.search(r"wavefront rays closest=(\d+) shadow=(\d+) transparent-shadow=(\d+) in (\d+) batches / (\d+) libb200rt calls \((\d+) rays per batch\), ([0-9.e+-]+) thread-seconds inside libb200rt of ([0-9.e+-]+) thread-seconds in the render workers; per-ray calls outside fibers: (\d+)", p.stdout)
        if wf:
            rec["wavefront"] = dict(zip(("closest", "shadow", "tshadow", "batches", "calls", "rays_per_batch"), map(int, wf.groups()[:6])), trace_thread_seconds=float(wf.group(7)), worker_thread_seconds=float(wf.group(8)), per_ray_calls=int(wf.group(9)))
            # all passes of the frame: kernel launches (the flush combiner merges the render threads' flushes) and rays
            lines = re.findall(r"wavefront rays closest=(\d+) shadow=(\d+) transparent-shadow=(\d+) .*?kernel launches: (\d+)", p.stdout)
            if lines:
                rec["wavefront"]["kernel_launches"] = sum(int(x[3]) for x in lines)
                rec["wavefront"]["rays_all_passes"] = sum(int(x[0]) + int(x[1]) + int(x[2]) for x in lines)
                rec["wavefront"]["rays_per_launch"] = rec["wavefront"]["rays_all_passes"] // max(1, rec["wavefront"]["kernel_launches"])
        if os.path.exists(out):
            img = read_tga(out)
            if first is None:
                first = img
            else:
                rec["psnr_vs_first_stock_db"] = psnr(first, img)
        print(json.dumps(rec), flush=True)

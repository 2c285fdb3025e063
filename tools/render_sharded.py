#!/usr/bin/env python
"""Tile-sharded `yafaray_render` over N GPUs of one box (BASELINE.json configs[2]).

One process per GPU (torchrun).  Every rank runs integration/_build/render_bench -- a client of libYafaRay's public C API
linked against the patched library -- on the same synthetic scene with the "b200-kdtree" accelerator on ITS GPU
(accelerator parameter device = LOCAL_RANK) and renders only its share of the frame's tiles (tile_shard_index / _count,
integration/include/render/tile_shard_b200.h).  Each rank's film (weighted sums, the reference's own ".film" file) is then
summed onto rank 0 with one NCCL reduce (libyafaray_b200/film.py) and normalised.  Rank 0 prints ONE JSON line:

    render_seconds      max over ranks of the time inside yafaray_render (scene build / kd build are per rank and not sharded)
    reduce_seconds      the film collective incl. host<->device copies
    mrays_per_s         rays traced by all ranks / render_seconds (from the wavefront queue's own counters)
    psnr_vs_single_db   with --compare: PSNR of the summed film against ONE process rendering every tile (linear, clipped to [0,1])

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \\
        tools/render_sharded.py --integrator pathtracing --width 1920 --height 1080 --aa 64 --compare
    python tools/render_sharded.py ...            (1 GPU)
    --accelerator yafaray-kdtree-original --backend gloo   runs the same thing on the CPU kd-tree (what tests/test_film.py does)
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BINARY = os.path.join(ROOT, "integration", "_build", "render_bench")
WF = re.compile(r"wavefront rays closest=(\d+) shadow=(\d+) transparent-shadow=(\d+) in (\d+) batches")


def render(args, out_dir, name, threads, extra):
    """One render_bench process; returns (its RENDER_BENCH record, rays traced, Film)."""
    from libyafaray_b200 import film
    prefix = os.path.join(out_dir, name)
    cmd = [BINARY, args.accelerator, args.integrator, str(args.cells), str(args.width), str(args.height), str(args.aa),
           prefix + ".tga", str(threads), "film_save=" + prefix] + extra + args.extra.split()
    p = subprocess.run(cmd, cwd=out_dir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, errors="replace")
    m = re.search(r"RENDER_BENCH (\{.*\})", p.stdout)
    if p.returncode != 0 or not m:
        raise RuntimeError(f"render_bench failed (exit {p.returncode}): {' '.join(cmd)}\n{p.stdout[-3000:]}")
    if args.accelerator == "b200-kdtree" and "no usable accelerator" in p.stdout:
        raise RuntimeError("libb200rt could not build the scene on this box (no CPU fallback):\n" + p.stdout[-2000:])
    rays = sum(sum(int(x) for x in w.groups()[:3]) for w in WF.finditer(p.stdout))
    return json.loads(m.group(1)), rays, film.read_film(film.film_path(prefix))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--accelerator", default="b200-kdtree")
    ap.add_argument("--integrator", default="pathtracing")
    ap.add_argument("--cells", type=int, default=707)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--aa", type=int, default=16)
    ap.add_argument("--threads", type=int, default=0, help="render threads per rank; 0 = host cores / ranks on this box")
    ap.add_argument("--backend", default="", help="nccl (default with b200-kdtree) or gloo")
    ap.add_argument("--compare", action="store_true", help="rank 0 also renders every tile in one process and reports the PSNR")
    ap.add_argument("--extra", default="", help="extra render_bench arguments, e.g. 'wavefront_fibers=1024 i:bounces=5'")
    ap.add_argument("--save", default="", help="rank 0 writes the summed film here (the reference's .film format)")
    ap.add_argument("--image", default="", help="rank 0 writes the normalised first layer here as a 32-bit TGA (sRGB)")
    args = ap.parse_args()

    import numpy as np
    import torch
    import torch.distributed as dist
    from libyafaray_b200 import film

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    backend = args.backend or ("nccl" if args.accelerator == "b200-kdtree" else "gloo")
    device = None
    if backend == "nccl":
        if not torch.cuda.is_available():
            raise SystemExit("render_sharded.py: the nccl backend needs CUDA devices")
        torch.cuda.set_device(local)
        device = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group(backend, rank=rank, world_size=world, **({"device_id": device} if device is not None else {}))
    threads = args.threads or max(1, (os.cpu_count() or 1) // local_world)
    extra = [f"tile_shard={rank}/{world}"]
    if args.accelerator == "b200-kdtree":
        extra.append(f"device={local}")

    # the ranks build the SAME kd-tree on host cores they share: the first to take the cache lock builds it, the others read it
    # (libyafaray_b200/csrc/b200rt.cu, "Tree cache")
    cache_dir = ""
    if world > 1 and args.accelerator == "b200-kdtree" and "B200RT_TREE_CACHE_DIR" not in os.environ:
        cache_dir = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir(), "b200rt_tree_cache_" + os.environ.get("MASTER_PORT", "0"))
        os.makedirs(cache_dir, exist_ok=True)
        os.environ["B200RT_TREE_CACHE_DIR"] = cache_dir

    with tempfile.TemporaryDirectory() as d:
        dist.barrier()
        w0 = time.perf_counter()
        rec, rays, mine = render(args, d, f"shard{rank}", threads, extra)
        total, reduce_s = film.reduce_film(mine, dst=0, device=device)
        wall = time.perf_counter() - w0
        stats = torch.tensor([rec["render_seconds"], rec["preprocess_seconds"], reduce_s, wall], dtype=torch.float64, device=device)
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        count = torch.tensor([rays, int((mine.weights > 0).sum())], dtype=torch.int64, device=device)
        dist.all_reduce(count, op=dist.ReduceOp.SUM)
        if rank == 0:
            render_s, pre_s, red_s, wall_s = [float(x) for x in stats.cpu()]
            line = {"tool": "render_sharded", "accelerator": args.accelerator, "integrator": args.integrator, "n_gpus": world if backend == "nccl" else 0,
                    "processes": world, "backend": backend, "threads_per_process": threads, "triangles": rec["triangles"],
                    "width": args.width, "height": args.height, "aa_samples": args.aa,
                    "render_seconds": render_s, "preprocess_seconds": pre_s, "reduce_seconds": red_s, "wall_seconds": wall_s,
                    "film_bytes_reduced": int(mine.flat().nbytes), "rays": int(count[0]),
                    "mrays_per_s": (int(count[0]) / render_s / 1e6) if render_s > 0 else None,
                    "pixels_with_weight": int((total.weights > 0).sum()), "pixels": args.width * args.height,
                    "shard_pixels_with_weight_sum": int(count[1])}
            if args.save:
                film.write_film(args.save, total)
            if args.image:
                film.write_tga(args.image, total)
            if args.compare:
                single_extra = ["tile_shard=0/1"] + ([f"device={local}"] if args.accelerator == "b200-kdtree" else [])
                rec1, rays1, single = render(args, d, "single", args.threads or (os.cpu_count() or 1), single_extra)
                line["single_render_seconds"] = rec1["render_seconds"]
                line["single_rays"] = rays1
                line["psnr_vs_single_db"] = film.psnr(film.normalized(total), film.normalized(single))
                # the noise floor of that comparison: the same single-process render twice (tile RNGs are seeded from rand()
                # and thread timing, SURVEY.md section 4, so two runs of a Monte Carlo integrator never agree exactly)
                _, _, again = render(args, d, "single2", args.threads or (os.cpu_count() or 1), single_extra)
                line["psnr_single_vs_single_db"] = film.psnr(film.normalized(again), film.normalized(single))
                line["max_weight_difference"] = float(np.abs(total.weights - single.weights).max())
            print(json.dumps(line), flush=True)
        dist.barrier()
    if cache_dir and rank == 0:
        import shutil
        shutil.rmtree(cache_dir, ignore_errors=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

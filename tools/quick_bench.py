#!/usr/bin/env python
"""Kernel-only timing of the device-resident closest + shadow passes (tuning aid; bench.py is the judged number).
   B200RT_LIB=build/variant.so python tools/quick_bench.py [--scene hf|obj|soup] [--rays N]"""
import argparse, hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from libyafaray_b200 import rt, scenes

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="hf")
ap.add_argument("--rays", type=int, default=1 << 24)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--leaf", type=int, default=0)
ap.add_argument("--cost", type=float, default=0.0)
ap.add_argument("--bonus", type=float, default=0.0)
ap.add_argument("--tag", default="")
ap.add_argument("--sort", default="", help="reorder the rays before tracing: 'morton' (origin cell, 10 bits per axis), 'octmorton' (direction octant, then origin), 'dirmorton' (origin 7 bits/axis + direction 3 bits/axis interleaved)")
a = ap.parse_args()
mesh = {"hf": lambda: scenes.heightfield(707), "obj": lambda: scenes.objects(1_000_000), "soup": lambda: scenes.soup(1_000_000),
        "obj10": lambda: scenes.objects(10_000_000)}[a.scene]()
sc = rt.Scene(0, rt.make_params(max_leaf_size=a.leaf, cost_ratio=a.cost, empty_bonus=a.bonus))
sc.add_mesh(*mesh); sc.build()
st = sc.stats()
n = a.rays
b = sc.bound()
rays = scenes.rays_incoherent(n, seed=12345, lo=b[:3] if a.scene != "hf" else (0, 0, 0), hi=b[3:] if a.scene != "hf" else (1, 1, 1))
def part1by2(x):
    x = x.astype(np.uint64) & 0x3FF
    x = (x | (x << 16)) & 0x30000FF
    x = (x | (x << 8)) & 0x300F00F
    x = (x | (x << 4)) & 0x30C30C3
    x = (x | (x << 2)) & 0x9249249
    return x
if a.sort:
    lo = np.asarray(b[:3], np.float64); ext = np.maximum(np.asarray(b[3:], np.float64) - lo, 1e-30)
    q = np.clip(((rays[:, 0:3] - lo) / ext * 1024).astype(np.int64), 0, 1023)
    key = part1by2(q[:, 0]) | (part1by2(q[:, 1]) << 1) | (part1by2(q[:, 2]) << 2)
    if a.sort == "octmorton":
        octant = ((rays[:, 4] < 0).astype(np.uint64)) | ((rays[:, 5] < 0).astype(np.uint64) << 1) | ((rays[:, 6] < 0).astype(np.uint64) << 2)
        key = key | (octant << 30)
    elif a.sort == "dirmorton":
        d = rays[:, 4:7] / np.linalg.norm(rays[:, 4:7], axis=1, keepdims=True)
        dq = np.clip(((d + 1) * 4).astype(np.int64), 0, 7)
        dkey = part1by2(dq[:, 0]) | (part1by2(dq[:, 1]) << 1) | (part1by2(dq[:, 2]) << 2)
        key = ((key >> 9) << 9) | dkey  # 7 bits per axis of origin, then 3 bits per axis of direction
    rays = np.ascontiguousarray(rays[np.argsort(key, kind="stable")])
srays = rays.copy(); srays[:, 3] = 0.0005; srays[:, 7] = 0.25
d_r = torch.from_numpy(rays).cuda(); d_s = torch.from_numpy(srays).cuda()
d_h = torch.empty((n, 4), dtype=torch.float32, device="cuda"); d_o = torch.empty(n, dtype=torch.int32, device="cuda")
sp = torch.cuda.current_stream().cuda_stream
def run(k):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tc = ts = 0.0
    for _ in range(k):
        ev[0].record(); sc.trace_closest_device(d_r.data_ptr(), n, d_h.data_ptr(), sp)
        ev[1].record(); sc.trace_shadow_device(d_s.data_ptr(), n, d_o.data_ptr(), sp)
        ev[2].record(); torch.cuda.synchronize()
        tc += ev[0].elapsed_time(ev[1]); ts += ev[1].elapsed_time(ev[2])
    return tc / k, ts / k
run(3)
tc, ts = run(a.steps)
h = hashlib.md5(d_h.cpu().numpy().tobytes()).hexdigest()[:8]
hs = hashlib.md5((d_o.cpu().numpy() != -1).tobytes()).hexdigest()[:8]
print(f"{a.tag or os.environ.get('B200RT_LIB', 'default'):40s} {a.scene} closest {n / tc / 1e3:8.1f} Mrays/s ({tc:.3f} ms)  shadow {n / ts / 1e3:8.1f} Mrays/s ({ts:.3f} ms)  "
      f"nodes {st['n_nodes']} refs {st['n_leaf_refs']} depth {st['max_depth']} build {st['build_seconds']:.2f}s  md5 {h} {hs}", flush=True)

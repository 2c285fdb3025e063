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
a = ap.parse_args()
mesh = {"hf": lambda: scenes.heightfield(707), "obj": lambda: scenes.objects(1_000_000), "soup": lambda: scenes.soup(1_000_000)}[a.scene]()
sc = rt.Scene(0, rt.make_params(max_leaf_size=a.leaf, cost_ratio=a.cost, empty_bonus=a.bonus))
sc.add_mesh(*mesh); sc.build()
st = sc.stats()
n = a.rays
b = sc.bound()
rays = scenes.rays_incoherent(n, seed=12345, lo=b[:3] if a.scene != "hf" else (0, 0, 0), hi=b[3:] if a.scene != "hf" else (1, 1, 1))
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

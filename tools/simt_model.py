#!/usr/bin/env python
"""Driver of tools/simt_model.cc (design aid): exports the S1M-hf tree libb200rt's builder makes plus a ray sample, builds and runs
the warp-level cost model.   python tools/simt_model.py [n_rays]"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from libyafaray_b200 import rt, scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
out = "/tmp/simt"
os.makedirs(out, exist_ok=True)
xyz, idx, flags = scenes.heightfield(707)
t = rt.host_tree(xyz, idx)
t["a"].tofile(f"{out}/a.bin"); t["b"].tofile(f"{out}/b.bin"); t["refs"].astype(np.uint32).tofile(f"{out}/refs.bin")
np.asarray(t["bound"], np.float32).tofile(f"{out}/bound.bin")
np.ascontiguousarray(xyz, np.float32).tofile(f"{out}/xyz.bin"); np.ascontiguousarray(idx, np.uint32).tofile(f"{out}/idx.bin")
scenes.rays_incoherent(n, seed=12345).tofile(f"{out}/rays.bin")
scenes.rays_shadow(n, seed=12346, t_max=0.25).tofile(f"{out}/srays.bin")
root = os.path.dirname(os.path.abspath(__file__))
subprocess.check_call(["g++", "-O2", "-o", f"{out}/simt_model", os.path.join(root, "simt_model.cc")])
for kind in ("closest", "shadow"):
    subprocess.check_call([f"{out}/simt_model", out, kind])

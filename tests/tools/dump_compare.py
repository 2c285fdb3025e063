#!/usr/bin/env python
"""Differential check on a scene dumped by AcceleratorB200 (B200_DUMP_SCENE=<file>, integration/src/accelerator/accelerator_b200.cc):
the flattened world-space geometry is traced by (a) the unmodified reference on its own kd-tree (oracle/_ref), (b) the oracle's
restatement of the reference traversal on libb200rt's HOST tree, and -- on a GPU box -- (c) libb200rt itself.  Primary rays come
from points around the scene, secondary rays start ON surfaces (hit point = from + t * dir, tmin = 0.0005) like the integrators'
bounce / refraction / shadow rays.  TEST INFRASTRUCTURE (uses oracle/).

    python tests/tools/dump_compare.py scene.bin [--rays 400000] [--save mismatches.npz]
"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from libyafaray_b200 import rt, scenes
from oracle import kdo, yref
from tests.helpers import host_tree_as_oracle_tree


from tests.helpers import load_scene_dump as load_dump, surface_rays


def secondary(rays, t, n, seed):
    return surface_rays(rays, t, seed)[:n]


def report(name, got, ref):
    same = got["prim"] == ref["prim"]
    hm = (got["prim"] >= 0) != (ref["prim"] >= 0)
    tdiff = same & (got["t"] != ref["t"])
    rel = np.abs(got["t"] - ref["t"]) / np.maximum(np.abs(ref["t"]), 1e-30)
    bad = (~same) & ((rel > 1e-5) | hm)
    print(f"{name}: id agreement {same.mean():.6f}, hit/miss disagreements {int(hm.sum())}, same id but different t {int(tdiff.sum())}, non-tie mismatches {int(bad.sum())}")
    return bad


ap = argparse.ArgumentParser()
ap.add_argument("dump")
ap.add_argument("--rays", type=int, default=400000)
ap.add_argument("--save", default="")
a = ap.parse_args()
xyz, idx, flags = load_dump(a.dump)
print("faces", idx.shape[0], "verts", xyz.shape[0], "flag values", np.unique(flags))
ref = yref.RefScene(xyz, idx, flags)
b = ref.bound()
ncpu = os.cpu_count() or 1
# primary rays from a shell around the geometry's median region
ctr = np.median(xyz, axis=0); ext = np.percentile(np.abs(xyz - ctr), 90, axis=0) + 1e-3
prim_rays = scenes.rays_incoherent(a.rays, seed=1, lo=ctr - 1.5 * ext, hi=ctr + 1.5 * ext)
r1 = ref.trace_closest(prim_rays, threads=ncpu)
sec_rays = secondary(prim_rays, r1["t"], a.rays, seed=2)
r2 = ref.trace_closest(sec_rays, threads=ncpu)
print(f"primary hits {np.mean(r1['prim'] >= 0):.3f}, secondary hits {np.mean(r2['prim'] >= 0):.3f}")
t = rt.host_tree(xyz, idx)
o = kdo.Oracle(xyz, idx, flags, tree=host_tree_as_oracle_tree(t), bound=t["bound"])
print("bound equal:", np.array_equal(t["bound"], b))
bad1 = report("reference traversal on libb200rt's host tree, primary", o.trace_closest(prim_rays, threads=ncpu), r1)
bad2 = report("reference traversal on libb200rt's host tree, secondary", o.trace_closest(sec_rays, threads=ncpu), r2)
sh_ref = ref.trace_shadow(sec_rays, threads=ncpu)["shadowed"]
print("shadow booleans differ (host tree):", int((o.trace_shadow(sec_rays, threads=ncpu)["shadowed"] != sh_ref).sum()))
if rt.device_count() > 0:
    s = rt.Scene(0); s.add_mesh(xyz, idx, flags); s.build()
    def gpu(rays):
        h = s.trace_closest(rays); p = h["prim"].astype(np.int64); p[p == 0xFFFFFFFF] = -1
        return dict(prim=p, t=h["t"], u=h["u"], v=h["v"])
    g1, g2 = gpu(prim_rays), gpu(sec_rays)
    bad1 = report("libb200rt on the GPU, primary", g1, r1)
    bad2 = report("libb200rt on the GPU, secondary", g2, r2)
    print("shadow booleans differ (GPU):", int(((s.trace_shadow(sec_rays) != 0xFFFFFFFF).astype(np.uint8) != sh_ref).sum()))
    if a.save:
        np.savez_compressed(a.save, rays=sec_rays[bad2], ref_prim=r2["prim"][bad2], ref_t=r2["t"][bad2], gpu_prim=g2["prim"][bad2], gpu_t=g2["t"][bad2],
                            rays1=prim_rays[bad1], ref_prim1=r1["prim"][bad1], ref_t1=r1["t"][bad1], gpu_prim1=g1["prim"][bad1], gpu_t1=g1["t"][bad1])

#!/usr/bin/env python
"""Kernel timing + parity sample in the regimes bench.py does not cover (tuning / DESIGN.md evidence; bench.py is the
judged number): the HBM-resident 10 M-triangle scene of BASELINE.json configs[3], S1M-obj and S1M-soup.

    python tests/tools/regime_bench.py --scene obj --tris 10000000 [--rays 16777216] [--check 200000]

Parity sample: the first --check rays are also traced by the oracle's C restatement ON THE TREE libb200rt's host
builder produced (exported through b200rt_host_tree_*), so a 10 M-triangle check needs no second build."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from libyafaray_b200 import rt, scenes

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="obj", choices=["hf", "obj", "soup"])
ap.add_argument("--tris", type=int, default=10_000_000)
ap.add_argument("--rays", type=int, default=1 << 24)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--check", type=int, default=200_000)
a = ap.parse_args()

t0 = time.time()
if a.scene == "hf":
    mesh = scenes.heightfield(int(round((a.tris / 2) ** 0.5)))
elif a.scene == "obj":
    mesh = scenes.objects(a.tris)
else:
    mesh = scenes.soup(a.tris)
xyz, idx, flags = mesh
sc = rt.Scene(0)
sc.add_mesh(xyz, idx, flags)
sc.build()
st = sc.stats()
b = sc.bound()
n = a.rays
rays = scenes.rays_incoherent(n, seed=12345, lo=b[:3], hi=b[3:])
srays = scenes.rays_shadow(n, seed=12346, lo=b[:3], hi=b[3:], t_max=0.25)
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
hits = d_h.cpu().numpy(); occ = d_o.cpu().numpy()
res = {"scene": a.scene, "faces": int(idx.shape[0]), "rays": n, "closest_mrays": n / tc / 1e3, "shadow_mrays": n / ts / 1e3, "closest_ms": tc, "shadow_ms": ts,
       "hit_fraction": float(np.mean(hits[:, 3].view(np.uint32) != rt.MISS)), "shadowed_fraction": float(np.mean(occ != -1)),
       "tree": {k: st[k] for k in ("n_nodes", "n_leaf_refs", "max_depth", "device_bytes", "build_seconds")}}
if a.check > 0:
    from oracle import kdo
    from tests import helpers
    m = min(a.check, n)
    t = rt.host_tree(xyz, idx)
    o = kdo.Oracle(xyz, idx, flags, tree=helpers.host_tree_as_oracle_tree(t), bound=t["bound"])
    ref = o.trace_closest(rays[:m], threads=os.cpu_count() or 1, counters=True)
    counters = ref["counters"]
    prim = hits[:m, 3].view(np.uint32).astype(np.int64); prim[prim == rt.MISS] = -1
    same = prim == ref["prim"]
    res["check"] = {"rays": m, "id_agreement": float(np.mean(same)), "t_bit_identical_where_same_id": bool(np.array_equal(hits[:m, 0][same], ref["t"][same])),
                    "max_rel_t_err_on_id_mismatch": float(np.max(np.abs(hits[:m, 0][~same] - ref["t"][~same]) / np.maximum(ref["t"][~same], 1e-30))) if np.any(~same) else 0.0}
    rs = o.trace_shadow(srays[:m], threads=os.cpu_count() or 1)
    res["check"]["shadow_booleans_identical"] = bool(np.array_equal((occ[:m] != -1).astype(np.uint8), rs["shadowed"]))
    if counters is not None:
        res["check"]["oracle_visits_per_ray_on_this_tree"] = counters.per_ray()
        res["check"]["algorithmic_bytes_per_closest_ray"] = counters.bytes_per_ray(48)
res["total_seconds"] = time.time() - t0
print(json.dumps(res), flush=True)

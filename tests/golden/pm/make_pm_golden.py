"""Generates tests/golden/pm/pm_*.npz from the UNMODIFIED reference (oracle/_ref/libyafref.so, built by `make -C oracle ref`
from /root/reference): the reference's own point kd-tree node array, PhotonMap::gather and PhotonMap::findNearest results
for small photon sets.  Run from the repo root:  python tests/golden/pm/make_pm_golden.py
The fixtures travel to the GPU box; /root/reference does not.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)
from libyafaray_b200 import scenes  # noqa: E402
from oracle import pmo  # noqa: E402

CASES = [("surfaces", 6000, 11), ("clusters", 5000, 12), ("lattice", 3000, 13), ("uniform", 37, 14)]
GATHERS = [(1, 1e-3), (8, 4e-3), (50, 2.5e-2), (20, 1e30)]
NEAREST = [1e-4, 1e-2, 1.0]


def main():
    assert pmo.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    for kind, n, seed in CASES:
        pos, dirs = scenes.photon_cloud(kind, n, seed)
        points, normals = scenes.gather_points(pos, 500, seed=seed + 100)
        ref = pmo.RefMap(pos, dirs, build_threads=2)
        a, b = ref.tree()
        out = dict(pos=pos, dirs=dirs, points=points, normals=normals, tree_a=a, tree_b=b, gathers=np.array(GATHERS, np.float64), nearest=np.array(NEAREST, np.float64))
        for i, (k, r2) in enumerate(GATHERS):
            idx, d2, cnt, rad = ref.gather(points, k, r2)
            out[f"g{i}_idx"], out[f"g{i}_d2"], out[f"g{i}_n"], out[f"g{i}_r"] = idx, d2, cnt, rad
        # per-point radii (the SPPM call sites pass each hit point's own radius)
        radii = (np.random.default_rng(seed).random(len(points)).astype(np.float32) * 0.03) ** 2
        idx, d2, cnt, rad = ref.gather(points, 12, 0.0, radii)
        out.update(radii=radii, gr_idx=idx, gr_d2=d2, gr_n=cnt, gr_r=rad)
        for i, dist in enumerate(NEAREST):
            out[f"n{i}"] = ref.nearest(points, normals, dist)
        path = os.path.join(ROOT, "tests", "golden", "pm", f"pm_{kind}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path), "bytes; mean found", [float(out[f"g{i}_n"].mean()) for i in range(len(GATHERS))])


if __name__ == "__main__":
    main()

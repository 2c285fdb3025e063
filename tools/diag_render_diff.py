#!/usr/bin/env python
"""Where do two renders of a reference test scene differ?  (diagnostic for tests/test_render.py)
    python tools/diag_render_diff.py test01"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.test_render import read_tga, render, BUILD, psnr
test = sys.argv[1] if len(sys.argv) > 1 else "test01"
binary = os.path.join(BUILD, "yafaray_" + test)
imgs = {}
for name, accel, fibers in (("stock1", "", 0), ("stock2", "", 0), ("b200-per-ray", "b200-kdtree", 0), ("b200-fibers", "b200-kdtree", 1024)):
    with tempfile.TemporaryDirectory() as d:
        render(binary, d, accel, {"B200_AA_PASSES": "1", "B200_DETERMINISTIC": "1", "B200_WAVEFRONT_FIBERS": str(fibers)})
        out = [f for f in os.listdir(d) if f.endswith(".tga")]
        imgs[name] = read_tga(os.path.join(d, out[0]))
base = imgs["stock1"]
for name, img in imgs.items():
    diff = np.abs(img - base)
    ys, xs = np.nonzero(diff.max(axis=2))
    print(f"{name:14s} vs stock1: differing bytes {int((diff > 0).sum())}, differing pixels {len(ys)}, max |diff| {diff.max():.0f}, PSNR {psnr(img, base):.2f} dB, "
          f"histogram of |diff| {np.bincount(diff.astype(int).ravel())[:6].tolist()}, rows {ys.min() if len(ys) else '-'}..{ys.max() if len(ys) else '-'}, cols {xs.min() if len(xs) else '-'}..{xs.max() if len(xs) else '-'}")
d = np.abs(imgs["b200-per-ray"] - base).max(axis=2)
np.save("gpurun_out/diag_diff.npy", d.astype(np.uint8))
a, b = base, imgs["b200-fibers"]
ys, xs = np.nonzero(np.abs(a - b).max(axis=2))
print("channel histogram of differing bytes:", [(int((a[..., c] != b[..., c]).sum())) for c in range(a.shape[2])])
for k in range(0, len(ys), max(1, len(ys) // 12)):
    print((int(ys[k]), int(xs[k])), a[ys[k], xs[k]].astype(int).tolist(), b[ys[k], xs[k]].astype(int).tolist())
print("sign of (b200 - stock):", int(((b - a) > 0).sum()), int(((b - a) < 0).sum()))

"""Photon-map gather on one B200 (include/b200pm.h): throughput of b200pm_gather next to the reference's
PhotonMap::gather timed on the host cores, with parity over the CPU sample.

    python tools/pm_bench.py [--photons 1000000] [--points 1000000] [--k 100] [--sq-radius 2.5e-4] [--steps 5] [--cpu-seconds 5]

One JSON line.  `value` = gather points per second with points and results resident in HBM (CUDA events around `steps`
launches on the launching stream); `e2e` = the same through b200pm_gather on host buffers (copies inside the timed region);
`cpu_baseline` = the UNMODIFIED reference (oracle/_ref, kind "reference") or the C restatement (kind "port") on a bounded
sample of the same points, all host threads.  bench.py calls run() for its "photon_gather" key; the oracle is used here only
as the checker / the CPU arm.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KERNEL = "b200pm::pmLookupPhasedKernel<1,true> (k > 16: phased lookup, one pop per step, heaps in the result array; k <= 16: pmLookupKernel<0>, heaps in shared memory)"


def run(photons=1_000_000, points=1_000_000, k=100, sq_radius=2.5e-4, steps=5, cpu_seconds=5.0, device=0, kind="surfaces", cpu=True):
    import torch

    from libyafaray_b200 import pm, rt, scenes

    dev = torch.device("cuda", device)
    pos, dirs = scenes.photon_cloud(kind, photons, seed=99)
    pts, nrm = scenes.gather_points(pos, points, seed=98, jitter=0.002)
    m = pm.PhotonMap(pos, dirs, device=device)
    stats = m.stats()
    d_pts = torch.from_numpy(pts).to(dev)
    out = m.gather_device(d_pts, k, sq_radius)
    for _ in range(2):
        m.gather_device(d_pts, k, sq_radius, out=out)
    torch.cuda.synchronize(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    launches0 = rt.launch_count()
    ev[0].record()
    for _ in range(steps):
        m.gather_device(d_pts, k, sq_radius, out=out)
    ev[1].record()
    torch.cuda.synchronize(dev)
    ms = ev[0].elapsed_time(ev[1]) / steps
    launches = rt.launch_count() - launches0
    found_dev = out[0].cpu().numpy().view(np.uint32)
    n_found_dev = out[1].cpu().numpy().view(np.uint32)
    # findNearest on the same points
    d_nrm = torch.from_numpy(nrm).to(dev)
    near = m.find_nearest_device(d_pts, d_nrm, sq_radius)
    torch.cuda.synchronize(dev)
    ev[0].record()
    for _ in range(steps):
        m.find_nearest_device(d_pts, d_nrm, sq_radius, out=near)
    ev[1].record()
    torch.cuda.synchronize(dev)
    near_ms = ev[0].elapsed_time(ev[1]) / steps
    # end to end: host buffers in, host buffers out -- page-locked (b200rt_host_alloc) and, for comparison, pageable
    pin = [rt.PinnedBuffer((points, 3), np.float32), rt.PinnedBuffer((points, k), pm.FOUND_DTYPE), rt.PinnedBuffer(points, np.uint32), rt.PinnedBuffer(points, np.float32)]
    pin[0].array[:] = pts
    outs = (pin[1].array, pin[2].array, pin[3].array)
    m.gather(pin[0].array, k, sq_radius, out=outs)
    e2e_steps = max(2, min(steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        m.gather(pin[0].array, k, sq_radius, out=outs)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    found, n_found, radius_out = outs[0].copy(), outs[1].copy(), outs[2].copy()
    for b in pin:
        b.free()
    t0 = time.perf_counter()
    found_p, n_found_p, _ = m.gather(pts, k, sq_radius)
    pageable_s = time.perf_counter() - t0
    valid = np.arange(k)[None, :] < n_found[:, None]
    same = bool(np.array_equal(n_found, n_found_dev) and np.array_equal(found["photon"][valid], found_dev[:, :, 0][valid]))
    line = {
        "metric": "photon_gather_points_per_second", "value": points / (ms * 1e-3) / 1e6, "unit": "Mpoints/s", "ms": ms, "steps": steps,
        "kernel": KERNEL, "gpu_launches": int(launches), "dtype": "f32",
        "config": {"workload": f"{photons} photons ({kind}), {points} gather points, k={k}, sq_radius={sq_radius}", "inputs": "points and results resident in HBM"},
        "mean_found": float(n_found.mean()), "full_fraction": float((n_found == k).mean()),
        "find_nearest": {"value": points / (near_ms * 1e-3) / 1e6, "unit": "Mpoints/s", "ms": near_ms, "kernel": "b200pm::pmLookupKernel<2>"},
        "e2e": {"value": points / e2e_s / 1e6, "unit": "Mpoints/s", "h2d_bytes": int(pts.nbytes), "d2h_bytes": int(found.nbytes + n_found.nbytes + radius_out.nbytes),
                "path": "b200pm_gather on page-locked host buffers (chunks over two streams: H2D, kernel, D2H overlap)", "same_as_device_resident": same,
                "pageable_value": points / pageable_s / 1e6, "pageable_same": bool(np.array_equal(n_found_p, n_found))},
        "tuning": {key: os.environ.get(key) for key in ("B200PM_KERNEL", "B200PM_ROUND", "B200PM_SMEM_K", "B200PM_PATIENCE") if os.environ.get(key) is not None},
        "tree": stats,
    }
    if cpu:
        from oracle import pmo

        threads = os.cpu_count() or 1
        use_ref = pmo.ref_available()
        arm = pmo.RefMap(pos, dirs, build_threads=threads, query_threads=threads) if use_ref else pmo.OracleMap(pos, dirs)
        sample = 20_000
        t0 = time.perf_counter()
        res = arm.gather(pts[:sample], k, sq_radius)
        dt = arm.seconds if use_ref else time.perf_counter() - t0
        if use_ref:
            # grow the sample towards the time budget (the reference arm is threaded, the port is not)
            sample = int(min(points, max(sample, sample * cpu_seconds / max(dt, 1e-6))))
            res = arm.gather(pts[:sample], k, sq_radius)
            dt = arm.seconds
        ok_valid = np.arange(k)[None, :] < res[2][:, None]
        parity = {"points": int(sample), "n_found_equal": bool(np.array_equal(res[2], n_found[:sample])),
                  "photon_order_mismatches": int(np.count_nonzero(res[0][ok_valid] != found["photon"][:sample][ok_valid])),
                  "dist_bits_mismatches": int(np.count_nonzero(res[1].view(np.uint32)[ok_valid] != found["dist_square"][:sample].view(np.uint32)[ok_valid])),
                  "radius_bits_mismatches": int(np.count_nonzero(res[3].view(np.uint32) != radius_out[:sample].view(np.uint32)))}
        line["cpu_baseline"] = {"value": sample / dt / 1e6, "unit": "Mpoints/s", "cores": threads if use_ref else 1, "kind": "reference" if use_ref else "port",
                                "sample": f"first {sample} of {points} points", "build_seconds": getattr(arm, "build_seconds", None)}
        line["parity"] = parity
    m.close()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--photons", type=int, default=1_000_000)
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--sq-radius", type=float, default=2.5e-4)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=5.0)
    ap.add_argument("--kind", default="surfaces")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    print(json.dumps(run(a.photons, a.points, a.k, a.sq_radius, a.steps, a.cpu_seconds, kind=a.kind, cpu=not a.no_cpu)), flush=True)


if __name__ == "__main__":
    main()

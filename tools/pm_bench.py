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


def gather_roofline(ms, photons, points, k, kind, sq_radius, tuning):
    """Roofline of the gather kernel on the DEFAULT workload: both bounds, the tighter one named; None for any
    other workload or tuning.  The counters are those of the committed ncu capture of this kernel
    (profiles/r6c_pm_gather_phased_single_pop.txt: one launch, patience 8; the default patience 16 is 3 % faster); `ms` is the
    duration measured live by the caller."""
    if (photons, points, k, kind) != (1_000_000, 1_000_000, 100, "surfaces") or abs(sq_radius - 2.5e-4) > 1e-12 or tuning:
        return None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    sm_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    warp_inst, dram_bytes = 4_538_847_914, 4_170_942_000 + 2_924_250_000
    algo_bytes = points * (12 + 8 * k + 8)
    issue = {"achieved": warp_inst / (ms * 1e-3) / 1e9, "peak": 148 * 4 * sm_mhz * 1e6 / 1e9, "unit": "G warp-inst/s", "warp_inst_per_point": warp_inst / points,
             "lanes_per_inst": 11.36, "peak_source": f"148 SMs x 4 schedulers x {sm_mhz:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz)"}
    issue["frac"] = issue["achieved"] / issue["peak"]
    hbm = {"achieved": algo_bytes / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "algorithmic_bytes_per_point": 12 + 8 * k + 8, "traffic": dram_bytes,
           "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"}
    hbm["frac"] = hbm["achieved"] / hbm["peak"]
    tighter = issue if issue["frac"] >= hbm["frac"] else hbm
    return {"bound": "issue" if tighter is issue else "hbm", "kernel": "b200pm::pmLookupPhasedKernel<1,true>", "launch_ms": ms, "achieved": tighter["achieved"],
            "peak": tighter["peak"], "unit": tighter["unit"], "frac": tighter["frac"], "traffic": dram_bytes, "issue": issue, "hbm": hbm,
            "counters_source": "profiles/r6c_pm_gather_phased_single_pop.txt (ncu --set full, one launch of this kernel on this workload)",
            "note": "tree (48 MB) and live heaps are L2 / L1 traffic; DRAM traffic is 8.6x the algorithmic bytes because the heaps of all resident warps do not fit L2 (DESIGN.md 11)"}


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
    roof = gather_roofline(ms, photons, points, k, kind, sq_radius, line["tuning"])
    if roof:
        line["roofline"] = roof
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

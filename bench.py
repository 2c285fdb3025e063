#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: Mrays/s closest-hit + shadow per B200 vs the host-CPU kd-tree.

Default workload (BASELINE.json configs[1], the judged line): synthetic 1 M-triangle scene (S1M-hf, 999 698 triangles,
SURVEY.md 8d), 16 Mi incoherent random rays; one "step" = one closest-hit pass over the 16 Mi rays (R-inc) plus one any-hit
shadow pass over 16 Mi shadow rays (R-shadow, t_max = 0.25) => 32 Mi rays per step per GPU.

    python bench.py [--gpus N --steps K --warmup W]            this framework (libb200rt, CUDA sm_100a)
    python bench.py --impl reference [...]                      the reference's CPU kd-tree on the host cores
    python bench.py --workload s10m | rcoh                      other regimes (not the judged line):
        s10m   config[3]'s geometry: 10 M-triangle object scene (2 GB on the device, HBM-resident), incoherent rays
        rcoh   S1M-hf, coherent rays: 1920x1080 pin-hole camera rays (jittered, 8 per pixel) + shadow rays from their hit
               points to one point light (SURVEY.md 8d R-coh, src/camera/camera_perspective.cc:165-180)

Prints ONE JSON line (rank 0).  Weak scaling: every rank traces its own rays against a replicated scene; there is no
collective on the data path, only a barrier and a MAX-reduce of the time.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mrays/s closest-hit+shadow per B200 (1/2/4/8 GPU) vs host-CPU kd-tree"
UNIT = "Mrays/s"
# ALGORITHMIC bytes per ray (DESIGN.md "Roofline"): IO + 8 B x (interior + leaf nodes visited) + 4 B x leaf refs
# + 36 B x triangle tests, visit counts taken from the REFERENCE's own kd-tree traversing this very workload
# (counting oracle, tests/tools/count_bytes.py).  IO = 32 B ray + 16 B hit (closest) / + 4 B (shadow).
ALGO_BYTES = {"s1m": {"closest": 345.0, "shadow": 230.8}}
SHADOW_TMAX = 0.25
TSHADOW_DEPTH = 4
# ncu counters of the kernels on each workload (dram bytes, warp instructions, lanes per instruction ...), written by
# `tools/ncu_summary.py json` from the committed capture and keyed by the hash of the kernel source they were taken from:
# a capture of an older kernel is reported as stale (traffic = null) instead of silently going out of date.
COUNTERS_JSON = os.path.join(ROOT, "profiles", "kernel_counters.json")
KERNEL_SOURCES = [os.path.join(ROOT, "libyafaray_b200", "csrc", f) for f in ("kd_kernels.cuh",)]
# batches of 32 Ki rays and more run as two passes: setupKernel<Q> (ray setup, bound misses answered) + traceKernel<Q,false,true> (queue-fed traversal)
KERNEL_NAMES = {"closest": "b200rt::setupKernel<0> + b200rt::traceKernel<0,false,true,false>", "shadow": "b200rt::setupKernel<1> + b200rt::traceKernel<1,false,true,false>",
                "tshadow": "b200rt::setupKernel<2> + b200rt::traceKernel<2,false,true,false>"}
SMS, SCHEDULERS_PER_SM = 148, 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="s1m", choices=["s1m", "s10m", "rcoh"])
    ap.add_argument("--rays", type=int, default=1 << 24, help="rays per query per GPU (16 Mi)")
    ap.add_argument("--cells", type=int, default=707, help="height-field cells per side (707 -> 999 698 triangles)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-tshadow", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="skip the photon-map gather leg (key photon_gather)")
    return ap.parse_args()


def kernel_source_hash():
    h = hashlib.sha256()
    for p in KERNEL_SOURCES:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def load_counters(workload):
    """(counters of this workload or None, note)."""
    try:
        rec = json.load(open(COUNTERS_JSON))
    except Exception as e:
        return None, f"no {os.path.relpath(COUNTERS_JSON, ROOT)} ({e.__class__.__name__})"
    if rec.get("kernel_source_sha") != kernel_source_hash():
        return None, f"stale: {os.path.relpath(COUNTERS_JSON, ROOT)} was captured from kernel source {rec.get('kernel_source_sha')}, this is {kernel_source_hash()}"
    w = rec.get("workloads", {}).get(workload)
    if not w:
        return None, f"no capture of workload {workload} in {os.path.relpath(COUNTERS_JSON, ROOT)}"
    return w, rec.get("source", "")


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md), through NVML
    (nvidia_ml_py) every few milliseconds -- a timed region of a few tens of ms is too short for `nvidia-smi -lms`."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80}

    def __init__(self, index: int, period_s: float = 0.004):
        self.rows, self.ok, self.stop_flag, self.period = [], False, False, period_s
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((time.perf_counter(), sm, mask))
            except Exception:
                pass
            time.sleep(self.period)

    def window(self, t0, t1):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "error": getattr(self, "err", "nvml unavailable")}
        for _ in range(100):  # a timed region shorter than one sampling period: wait for the first samples around it
            if len(self.rows) >= 2:
                break
            time.sleep(0.002)
        rows = [r for r in self.rows if t0 <= r[0] <= t1]
        if len(rows) < 2:  # region shorter than two periods: take the nearest samples around it
            rows = sorted(self.rows, key=lambda r: abs(r[0] - 0.5 * (t0 + t1)))[:4]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": [], "samples": 0}
        mask = 0
        for r in rows:
            mask |= r[2]
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": self.max_sm,
                "reasons": sorted(k for k, bit in self.REASONS.items() if mask & bit), "samples": len(rows)}

    def stop(self):
        self.stop_flag = True


def bind_to_gpu_numa_node(index):
    """Pin this rank (and the pinned host buffers it allocates from here on, first touch) to the NUMA node its GPU hangs off:
    at N = 8 every rank streams 1.3 GiB per step through the host's memory controllers, and a buffer on the other socket
    crosses the inter-socket link on its way to the PCIe root complex.  Returns a description for the bench line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        path = "/sys/bus/pci/devices/" + bus.lower()[-12:]
        node = int(open(path + "/numa_node").read())
        if node < 0:
            return {"numa_node": None, "note": "the platform reports no NUMA node for this GPU"}
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus_bound": len(allowed)}
    except Exception as e:  # not fatal: the run goes on unbound
        return {"numa_node": None, "note": f"not bound ({e.__class__.__name__})"}


# ------------------------------------------------------------------------------------------ workload
LIGHT = np.array([0.5, 0.5, 3.0])  # rcoh: the point light the shadow rays aim at


def make_scene_arrays(args):
    from libyafaray_b200 import scenes
    if args.workload == "s10m":
        return scenes.objects(10_000_000)
    return scenes.heightfield(args.cells)


def make_closest_rays(args, rank, bound):
    from libyafaray_b200 import scenes
    n = args.rays
    if args.workload == "s1m":
        return scenes.rays_incoherent(n, seed=12345 + 2 * rank)
    if args.workload == "s10m":
        return scenes.rays_incoherent(n, seed=12345 + 2 * rank, lo=bound[:3], hi=bound[3:])
    w, h = 1920, 1080
    frames = [scenes.rays_camera(w, h, seed=1000 * rank + k) for k in range((n + w * h - 1) // (w * h))]
    return np.ascontiguousarray(np.concatenate(frames)[:n])


def make_shadow_rays(args, rank, bound, closest_rays=None, closest_t=None):
    """s1m / s10m: R-shadow (independent incoherent rays with a finite t_max).  rcoh: from the hit point of every camera ray
    (the ray origin where it missed) to the light, direction unnormalised so that t = 1 is the light (as the reference's
    light sampling does: tmax = distance along a unit direction; the two are equivalent for the traversal)."""
    from libyafaray_b200 import scenes
    n = args.rays
    if args.workload == "s1m":
        return scenes.rays_shadow(n, seed=12346 + 2 * rank, t_max=SHADOW_TMAX)
    if args.workload == "s10m":
        diag = float(np.linalg.norm(bound[3:].astype(np.float64) - bound[:3].astype(np.float64)))
        return scenes.rays_shadow(n, seed=12346 + 2 * rank, t_max=SHADOW_TMAX * diag, lo=bound[:3], hi=bound[3:])
    o = closest_rays[:, 0:3].astype(np.float64) + closest_rays[:, 4:7].astype(np.float64) * closest_t.astype(np.float64)[:, None]
    r = np.empty((n, 8), np.float32)
    r[:, 0:3] = o
    r[:, 3] = 0.0005
    r[:, 4:7] = LIGHT[None, :] - o
    r[:, 7] = 1.0
    return r


def workload_name(args, n_faces):
    if args.workload == "s1m":
        return f"S1M-hf height-field {n_faces} triangles; {args.rays} incoherent closest-hit rays + {args.rays} shadow rays (t_max {SHADOW_TMAX}) per GPU per step"
    if args.workload == "s10m":
        return f"S10M objects {n_faces} faces (HBM-resident, config[3] geometry); {args.rays} incoherent closest-hit rays + {args.rays} shadow rays (t_max {SHADOW_TMAX} x diagonal) per GPU per step"
    return f"R-coh: S1M-hf {n_faces} triangles; {args.rays} jittered 1920x1080 camera rays + {args.rays} shadow rays from their hit points to a point light per GPU per step"


def config_dict(args, n_faces):
    """The SAME keys and values in both arms (the driver compares the two `config` objects)."""
    return {"workload": workload_name(args, n_faces), "rays_per_step_per_gpu": 2 * args.rays,
            "l2_policy": "inputs larger than L2: 512 MiB rays + 256 MiB hits (closest), 512 MiB + 64 MiB (shadow) streamed per step at the default size; "
                         "the scene is read through L2 as the traversal finds it"}


# ------------------------------------------------------------------------------------------ CPU arm
class CpuArm:
    """The reference's CPU kd-tree on the host cores: the UNMODIFIED reference (oracle/_ref/libyafref.so, built by
    oracle/Makefile from /root/reference) when that library travelled with the repo, else the oracle's C
    restatement (oracle/kd_oracle.c) on its own tree.  Timed on a bounded sample of the workload."""

    def __init__(self, xyz, idx, flags, threads):
        from oracle import kdo, yref
        self.threads = threads
        if yref.available():
            self.kind = "reference"
            self.obj = yref.RefScene(xyz, idx, flags)
            self.build_seconds = self.obj.build_seconds
            self.what = "unmodified reference (oracle/_ref/libyafref.so, yafaray-kdtree-original)"
        else:
            self.kind = "port"
            t0 = time.perf_counter()
            self.obj = kdo.Oracle(xyz, idx, flags)
            self.build_seconds = time.perf_counter() - t0
            self.what = "C restatement (oracle/kd_oracle.c) on its own SAH tree"
        self.last = None

    def _timed(self, fn, r):
        t0 = time.perf_counter()
        out = fn(r, threads=self.threads)
        dt = time.perf_counter() - t0
        return (out.get("seconds", dt) if self.kind == "reference" else dt), out

    def measure(self, rays, srays, n):
        tc, c = self._timed(self.obj.trace_closest, rays[:n])
        ts, s = self._timed(self.obj.trace_shadow, srays[:n])
        self.last = (n, c, s)
        return tc, ts

    def sample_size(self, rays, srays, target_seconds):
        probe = min(1 << 18, rays.shape[0])
        tc, ts = self.measure(rays, srays, probe)
        rate = 2 * probe / max(tc + ts, 1e-9)
        return int(min(rays.shape[0], max(probe, rate * target_seconds / 2)))

    def describe(self, n, total):
        return f"first {n} closest + first {n} shadow rays of the {total}-ray workload, {self.threads} host threads, {self.what}"


def parity_report(n, cpu_closest, cpu_shadow, gpu_hits, gpu_occ, miss):
    """GPU results against the CPU arm's over the rays the CPU sample covered (all of them at the default size)."""
    prim = gpu_hits["prim"][:n].astype(np.int64)
    prim[prim == miss] = -1
    ref_prim = cpu_closest["prim"][:n].astype(np.int64)
    same = prim == ref_prim
    diff = np.flatnonzero(~same)
    t_g, t_r = gpu_hits["t"][:n], cpu_closest["t"][:n]
    both_hit = (prim[diff] >= 0) & (ref_prim[diff] >= 0)
    tie = both_hit & (np.abs(t_g[diff].astype(np.float64) - t_r[diff]) <= 1e-5 * np.maximum(np.abs(t_r[diff]), 1e-30))
    bits = (np.array_equal(t_g[same], t_r[same]) and np.array_equal(gpu_hits["u"][:n][same], cpu_closest["u"][:n][same])
            and np.array_equal(gpu_hits["v"][:n][same], cpu_closest["v"][:n][same]))
    sh = (gpu_occ[:n] != miss).astype(np.uint8)
    return {"rays_compared": int(n), "ids_equal": float(same.mean()) if n else None, "id_mismatches": int(diff.size),
            "non_tie_mismatches": int(diff.size - tie.sum()), "tuv_bit_identical_where_ids_equal": bool(bits),
            "shadow_bool_mismatches": int(np.count_nonzero(sh != cpu_shadow["shadowed"][:n])),
            "against": "cpu_baseline arm (same rays, same run)"}


def run_reference(args):
    """--impl reference: rank 0 alone times the CPU kd-tree; every step is one bounded sample."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    xyz, idx, flags = make_scene_arrays(args)
    steps, warm = max(1, args.steps), max(0, args.warmup)
    arm = CpuArm(xyz, idx, flags, threads)
    bound = arm.obj.bound()
    rays = make_closest_rays(args, 0, bound)
    if args.workload == "rcoh":
        srays = make_shadow_rays(args, 0, bound, rays, arm.obj.trace_closest(rays, threads=threads)["t"])
    else:
        srays = make_shadow_rays(args, 0, bound)
    per_step = max(1.0, min(args.cpu_seconds, 120.0 / (steps + warm)))
    n = arm.sample_size(rays, srays, per_step)
    for _ in range(warm):
        arm.measure(rays, srays, n)
    tcs, tss = zip(*[arm.measure(rays, srays, n) for _ in range(steps)])
    tc, ts = float(np.sum(tcs)), float(np.sum(tss))
    value = 2 * n * steps / (tc + ts) / 1e6
    base = {"value": value, "unit": UNIT, "cores": threads, "kind": arm.kind, "sample": arm.describe(n, args.rays),
            "closest_mrays": n * steps / tc / 1e6, "shadow_mrays": n * steps / ts / 1e6, "build_seconds": arm.build_seconds,
            "sample_rays_per_step": 2 * n}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": (tc + ts) / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_dict(args, idx.shape[0]),
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def transparent_flags(n_faces):
    """Every third face a transparent shadow caster: the transparent-shadow pass then really collects casters."""
    from libyafaray_b200 import scenes
    fl = np.full(n_faces, scenes.F_NORMAL, np.uint8)
    fl[::3] |= scenes.F_TRANSPARENT
    return fl


def run_b200(args):
    import torch
    import torch.distributed as dist
    from libyafaray_b200 import rt

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or rt.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device -- libb200rt has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    # only when several ranks share the host: at N = 1 the cpu_baseline leg below wants every core of the box
    affinity = bind_to_gpu_numa_node(local) if world > 1 else {"numa_node": None, "note": "single rank: not bound"}
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    xyz, idx, flags = make_scene_arrays(args)
    scene = rt.Scene(local)
    scene.add_mesh(xyz, idx, flags)
    scene.build()
    stats = scene.stats()
    bound = scene.bound()
    n = args.rays
    rays = make_closest_rays(args, rank, bound)
    if args.workload == "rcoh":
        srays = make_shadow_rays(args, rank, bound, rays, scene.trace_closest(rays)["t"])  # set-up, not timed
    else:
        srays = make_shadow_rays(args, rank, bound)

    # ---- device-resident arm: inputs already in HBM when the timed region starts ----
    d_rays = torch.from_numpy(rays).to(dev)
    d_srays = torch.from_numpy(srays).to(dev)
    d_hits = torch.empty((n, 4), dtype=torch.float32, device=dev)
    d_occ = torch.empty(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    def step(events=None):
        if events is not None:
            events[0].record(stream)
        scene.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr(), sp)
        if events is not None:
            events[1].record(stream)
        scene.trace_shadow_device(d_srays.data_ptr(), n, d_occ.data_ptr(), sp)
        if events is not None:
            events[2].record(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    launches0 = rt.launch_count()
    barrier()
    w0 = time.perf_counter()
    for k in range(args.steps):
        step(evs[k])
    end = torch.cuda.Event(enable_timing=True)
    end.record(stream)
    barrier()
    w1 = time.perf_counter()
    launches = rt.launch_count() - launches0
    total_ms = evs[0][0].elapsed_time(end)
    closest_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    shadow_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    clocks = sampler.window(w0, w1) if sampler else None
    gpu_hits = d_hits.cpu().numpy().view(rt.HIT_DTYPE).reshape(-1)
    gpu_occ = d_occ.cpu().numpy().view(np.uint32)

    # ---- end-to-end arm: the host-buffer C-ABI call, pinned host rays in, host results out ----
    e2e_s = None
    if not args.no_e2e:
        pin_r = rt.PinnedBuffer((n, 8), np.float32); pin_r.array[:] = rays
        pin_s = rt.PinnedBuffer((n, 8), np.float32); pin_s.array[:] = srays
        pin_h = rt.PinnedBuffer((n,), rt.HIT_DTYPE)
        pin_o = rt.PinnedBuffer((n,), np.uint32)
        for _ in range(2):
            scene.trace_closest(pin_r.array, out=pin_h.array)
            scene.trace_shadow(pin_s.array, out=pin_o.array)
        barrier()
        e0 = time.perf_counter()
        e2e_steps = max(2, min(args.steps, 5))
        for _ in range(e2e_steps):
            scene.trace_closest(pin_r.array, out=pin_h.array)
            scene.trace_shadow(pin_s.array, out=pin_o.array)
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - e0) / e2e_steps
        # results of both arms must be the same bytes
        assert pin_h.array.tobytes() == gpu_hits.tobytes(), "e2e and device-resident results differ"
        del pin_r, pin_s, pin_h, pin_o

    # ---- transparent shadows (traceKernel<2,*>): own key, outside `value`; a second copy of the scene in which every third
    # face is a transparent caster, the step's shadow rays, max_depth 4 ----
    tshadow = None
    if not args.no_tshadow and rank == 0:
        ts_scene = rt.Scene(local)
        ts_scene.add_mesh(xyz, idx, transparent_flags(idx.shape[0]))
        ts_scene.build()
        d_ts = torch.empty((n, rt.TSHADOW_DTYPE.itemsize // 4), dtype=torch.int32, device=dev)
        for _ in range(2):
            ts_scene.trace_tshadow_device(d_srays.data_ptr(), n, TSHADOW_DEPTH, d_ts.data_ptr(), sp)
        t_ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        reps = max(2, min(args.steps, 5))
        t_ev[0].record(stream)
        for _ in range(reps):
            ts_scene.trace_tshadow_device(d_srays.data_ptr(), n, TSHADOW_DEPTH, d_ts.data_ptr(), sp)
        t_ev[1].record(stream)
        torch.cuda.synchronize()
        ts_ms = t_ev[0].elapsed_time(t_ev[1]) / reps
        res = d_ts.cpu().numpy().view(rt.TSHADOW_DTYPE).reshape(-1)
        tshadow = {"kernel": KERNEL_NAMES["tshadow"], "mrays_per_gpu": n / (ts_ms * 1e-3) / 1e6, "ms": ts_ms, "max_depth": TSHADOW_DEPTH,
                   "transparent_faces": "every third", "result_bytes_per_ray": rt.TSHADOW_DTYPE.itemsize,
                   "shadowed_fraction": float(np.mean(res["shadowed"] != 0)), "mean_transparent_casters_on_lit_rays": float(np.mean(res["n_transparent"][res["shadowed"] == 0]))}
        del d_ts
        ts_scene.close()

    # ---- photon-map gather (include/b200pm.h): own key, outside `value` and `gpu_launches`; tools/pm_bench.py ----
    photon_gather = None
    if not args.no_gather and rank == 0 and args.workload == "s1m":
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("pm_bench", os.path.join(ROOT, "tools", "pm_bench.py"))
            pm_bench = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(pm_bench)
            photon_gather = pm_bench.run(steps=max(2, min(args.steps, 5)), cpu_seconds=min(5.0, args.cpu_seconds), device=local,
                                         cpu=(world == 1 and not args.no_cpu_baseline))
        except Exception as exc:  # this leg must never take the judged line down
            photon_gather = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- max over ranks ----
    t = torch.tensor([total_ms, closest_ms, shadow_ms, (e2e_s or 0.0) * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, closest_ms, shadow_ms, e2e_ms = [float(x) for x in t.cpu()]
    if sampler:
        sampler.stop()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        ms_per_step = total_ms / args.steps
        value = world * 2 * n / (ms_per_step * 1e-3) / 1e6
        counters, counters_note = load_counters(args.workload)
        full_size = n == (1 << 24) and (args.workload != "s1m" or args.cells == 707)
        cc = (counters or {}).get("closest") if full_size else None
        if cc and "setup" in cc:
            # two-pass batch: closest_ms spans the setup pass and the traversal pass, so do the counters
            cc = dict(cc)
            for k in ("dram_bytes", "warp_inst", "l2_sector_bytes"):
                cc[k] = cc[k] + cc["setup"][k]
            cc["lanes_per_inst"] = None
        # HBM bound: algorithmic bytes of one closest launch / its measured duration
        algo = ALGO_BYTES.get(args.workload)
        hbm = None
        if algo:
            achieved = algo["closest"] * n / (closest_ms * 1e-3) / 1e9
            hbm = {"achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "peak_source": peak_src,
                   "algorithmic_bytes_per_ray": algo["closest"], "algorithmic_bytes_per_launch": algo["closest"] * n,
                   "traffic": cc["dram_bytes"] if cc else None}
        elif cc:
            # no counted visit numbers for this workload: the bytes the launch moved through L2 in 32-byte sectors (ncu) over the live time
            achieved = cc["l2_sector_bytes"] / (closest_ms * 1e-3) / 1e9
            hbm = {"achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "peak_source": peak_src,
                   "bytes": "L2 sector bytes of one closest launch (ncu lts__t_sectors x 32 B), not algorithmic", "traffic": cc["dram_bytes"]}
        # issue bound: warp instructions of one closest launch (ncu) / its duration, against SMs x 4 schedulers x SM clock
        issue = None
        sm_mhz = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
        if cc:
            peak_issue = SMS * SCHEDULERS_PER_SM * sm_mhz * 1e6 / 1e9
            ach_issue = cc["warp_inst"] / (closest_ms * 1e-3) / 1e9
            issue = {"achieved": ach_issue, "peak": peak_issue, "unit": "G warp-inst/s", "frac": ach_issue / peak_issue,
                     "warp_inst_per_ray": cc["warp_inst"] / n, "lanes_per_inst": cc.get("lanes_per_inst"),
                     "peak_source": f"{SMS} SMs x {SCHEDULERS_PER_SM} schedulers x {sm_mhz:.0f} MHz (clock sampled during the timed region)"}
        bounds = {k: v for k, v in (("hbm", hbm), ("issue", issue)) if v}
        tighter = max(bounds, key=lambda k: bounds[k]["frac"]) if bounds else None
        roof = {"bound": tighter, "kernel": KERNEL_NAMES["closest"], "launch_ms": closest_ms, "counters_source": counters_note}
        if tighter:
            roof.update({k: bounds[tighter][k] for k in ("achieved", "peak", "unit", "frac")})
            roof["traffic"] = (hbm or {}).get("traffic")
        roof.update(bounds)
        roof["note"] = ("a gather over an L2-resident scene: the algorithmic-byte fraction of HBM is small by construction; instruction issue is the "
                        "tighter bound (DESIGN.md 5)")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": config_dict(args, idx.shape[0]),
            "detail": {"closest_mrays_per_gpu": n / (closest_ms * 1e-3) / 1e6, "shadow_mrays_per_gpu": n / (shadow_ms * 1e-3) / 1e6,
                       "closest_ms": closest_ms, "shadow_ms": shadow_ms, "kernels": [KERNEL_NAMES["closest"], KERNEL_NAMES["shadow"]],
                       "kernel_source_sha": kernel_source_hash(),
                       "tree": {k: stats[k] for k in ("n_nodes", "n_leaf_refs", "max_depth", "device_bytes", "build_seconds")}},
            "roofline": roof,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "wall_ms_per_step": (w1 - w0) * 1e3 / args.steps,
            "host_affinity": affinity,
        }
        if tshadow:
            line["tshadow"] = tshadow
        if photon_gather:
            line["photon_gather"] = photon_gather
        if e2e_s is not None:
            line["e2e"] = {"value": world * 2 * n / (e2e_ms * 1e-3) / 1e6, "unit": UNIT,
                           "h2d_bytes_per_step": 2 * n * 32, "d2h_bytes_per_step": n * 16 + n * 4,
                           "path": "b200rt_trace_closest + b200rt_trace_shadow on pinned host buffers (H2D, kernel, D2H chunk-pipelined inside the call)"}
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            arm = CpuArm(xyz, idx, flags, threads)
            m = arm.sample_size(rays, srays, args.cpu_seconds)
            tc, ts = arm.measure(rays, srays, m)
            line["cpu_baseline"] = {"value": 2 * m / (tc + ts) / 1e6, "unit": UNIT, "cores": threads, "kind": arm.kind, "sample": arm.describe(m, n),
                                    "closest_mrays": m / tc / 1e6, "shadow_mrays": m / ts / 1e6, "build_seconds": arm.build_seconds}
            _, c_res, s_res = arm.last
            line["parity"] = parity_report(m, c_res, s_res, gpu_hits, gpu_occ, rt.MISS)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

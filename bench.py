#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: Mrays/s closest-hit + shadow per B200 vs the host-CPU kd-tree.

Workload (BASELINE.json configs[1]): synthetic 1 M-triangle scene (S1M-hf, 999 698 triangles),
16 M incoherent random rays; one "step" = one closest-hit pass over the 16 M rays (R-inc) plus one any-hit
shadow pass over 16 M shadow rays (R-shadow, t_max = 0.25) => 32 M rays per step per GPU.

    python bench.py [--gpus N --steps K --warmup W]            this framework (libb200rt, CUDA sm_100a)
    python bench.py --impl reference [...]                      the reference's CPU kd-tree on the host cores

Prints ONE JSON line (rank 0).  Weak scaling: every rank traces its own 16 M + 16 M rays against a replicated
scene; there is no collective on the data path, only a barrier and a MAX-reduce of the time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
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
ALGO_BYTES = {"closest": 345.0, "shadow": 230.8}
SHADOW_TMAX = 0.25
# dram__bytes_read.sum + dram__bytes_write.sum of ONE closest launch on this workload, from the ncu --set full capture kept
# under profiles/ (r2v_final_kernels.txt: 1.588 GB read + 0.292 GB written); a constant of the profile, not measured live
NCU_DRAM_TRAFFIC_BYTES = {"closest": 1.587585e9 + 0.292205e9, "source": "profiles/r2v_final_kernels.txt"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rays", type=int, default=1 << 24, help="rays per query per GPU (16 Mi)")
    ap.add_argument("--cells", type=int, default=707, help="height-field cells per side (707 -> 999 698 triangles)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


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


# ------------------------------------------------------------------------------------------ workload
def make_scene_arrays(cells):
    from libyafaray_b200 import scenes
    return scenes.heightfield(cells)


def make_rays(n, rank):
    from libyafaray_b200 import scenes
    return (scenes.rays_incoherent(n, seed=12345 + 2 * rank),
            scenes.rays_shadow(n, seed=12346 + 2 * rank, t_max=SHADOW_TMAX))


def workload_name(args, n_faces):
    return f"S1M-hf height-field {n_faces} triangles; {args.rays} incoherent closest-hit rays + {args.rays} shadow rays (t_max {SHADOW_TMAX}) per GPU per step"


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

    def _timed(self, fn, r):
        t0 = time.perf_counter()
        out = fn(r, threads=self.threads)
        dt = time.perf_counter() - t0
        return out.get("seconds", dt) if self.kind == "reference" else dt

    def measure(self, rays, srays, n):
        tc = self._timed(self.obj.trace_closest, rays[:n])
        ts = self._timed(self.obj.trace_shadow, srays[:n])
        return tc, ts

    def sample_size(self, rays, srays, target_seconds):
        probe = min(1 << 18, rays.shape[0])
        tc, ts = self.measure(rays, srays, probe)
        rate = 2 * probe / max(tc + ts, 1e-9)
        return int(min(rays.shape[0], max(probe, rate * target_seconds / 2)))

    def describe(self, n, total):
        return f"first {n} closest + first {n} shadow rays of the {total}-ray workload, {self.threads} host threads, {self.what}"


def cpu_baseline(xyz, idx, flags, rays, srays, target_seconds):
    threads = os.cpu_count() or 1
    arm = CpuArm(xyz, idx, flags, threads)
    n = arm.sample_size(rays, srays, target_seconds)
    tc, ts = arm.measure(rays, srays, n)
    return {"value": 2 * n / (tc + ts) / 1e6, "unit": UNIT, "cores": threads, "kind": arm.kind, "sample": arm.describe(n, rays.shape[0]),
            "closest_mrays": n / tc / 1e6, "shadow_mrays": n / ts / 1e6, "build_seconds": arm.build_seconds}


def run_reference(args):
    """--impl reference: rank 0 alone times the CPU kd-tree; every step is one bounded sample."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    xyz, idx, flags = make_scene_arrays(args.cells)
    rays, srays = make_rays(args.rays, 0)
    steps, warm = max(1, args.steps), max(0, args.warmup)
    arm = CpuArm(xyz, idx, flags, threads)
    per_step = max(1.0, min(args.cpu_seconds, 120.0 / (steps + warm)))
    n = arm.sample_size(rays, srays, per_step)
    for _ in range(warm):
        arm.measure(rays, srays, n)
    tcs, tss = zip(*[arm.measure(rays, srays, n) for _ in range(steps)])
    tc, ts = float(np.sum(tcs)), float(np.sum(tss))
    value = 2 * n * steps / (tc + ts) / 1e6
    base = {"value": value, "unit": UNIT, "cores": threads, "kind": arm.kind, "sample": arm.describe(n, args.rays),
            "closest_mrays": n * steps / tc / 1e6, "shadow_mrays": n * steps / ts / 1e6, "build_seconds": arm.build_seconds}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": (tc + ts) / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": workload_name(args, idx.shape[0]), "sample_rays_per_step": 2 * n},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    xyz, idx, flags = make_scene_arrays(args.cells)
    scene = rt.Scene(local)
    scene.add_mesh(xyz, idx, flags)
    scene.build()
    stats = scene.stats()
    rays, srays = make_rays(args.rays, rank)
    n = args.rays

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
        assert pin_h.array.tobytes() == d_hits.cpu().numpy().tobytes(), "e2e and device-resident results differ"

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
        achieved = ALGO_BYTES["closest"] * n / (closest_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args, idx.shape[0]), "rays_per_step_per_gpu": 2 * n,
                       "l2_policy": "inputs larger than L2: 512 MiB rays + 256 MiB hits (closest), 512 MiB + 64 MiB (shadow) streamed per step; "
                                    "the 1 M-triangle scene itself is L2-resident by nature of the workload",
                       "closest_mrays_per_gpu": n / (closest_ms * 1e-3) / 1e6, "shadow_mrays_per_gpu": n / (shadow_ms * 1e-3) / 1e6,
                       "closest_ms": closest_ms, "shadow_ms": shadow_ms, "tree": {k: stats[k] for k in ("n_nodes", "n_leaf_refs", "max_depth", "device_bytes", "build_seconds")}},
            "roofline": {"bound": "hbm", "kernel": "traceClosestKernel", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": NCU_DRAM_TRAFFIC_BYTES["closest"] if n == (1 << 24) and args.cells == 707 else None,
                         "traffic_source": NCU_DRAM_TRAFFIC_BYTES["source"], "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": ALGO_BYTES["closest"] * n,
                         "algorithmic_bytes_per_ray": ALGO_BYTES["closest"], "launch_ms": closest_ms,
                         "note": "latency/issue-bound gather (DESIGN.md): the algorithmic-byte fraction is small by construction"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "wall_ms_per_step": (w1 - w0) * 1e3 / args.steps,
        }
        if e2e_s is not None:
            line["e2e"] = {"value": world * 2 * n / (e2e_ms * 1e-3) / 1e6, "unit": UNIT,
                           "h2d_bytes_per_step": 2 * n * 32, "d2h_bytes_per_step": n * 16 + n * 4,
                           "path": "b200rt_trace_closest + b200rt_trace_shadow on pinned host buffers (H2D, kernel, D2H chunk-pipelined inside the call)"}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(xyz, idx, flags, rays, srays, args.cpu_seconds)
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

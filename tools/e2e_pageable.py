#!/usr/bin/env python
"""Host-buffer queries on PAGEABLE memory (what a caller with malloc'd / std::vector buffers has): Mrays/s of b200rt_trace_closest +
b200rt_trace_shadow on the bench workload, against the pinned-buffer rate.  B200RT_STAGING_HELPERS=0..3 sets the helper threads
of the staging copies (default 3 where the host has the cores)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from libyafaray_b200 import rt, scenes

n = 1 << 24
xyz, idx, flags = scenes.heightfield(707)
s = rt.Scene(0); s.add_mesh(xyz, idx, flags); s.build()
rays = scenes.rays_incoherent(n, seed=12345); srays = scenes.rays_shadow(n, seed=12346, t_max=0.25)
hits = np.empty(n, rt.HIT_DTYPE); occ = np.empty(n, np.uint32)
pr = rt.PinnedBuffer((n, 8), np.float32); pr.array[:] = rays
ps = rt.PinnedBuffer((n, 8), np.float32); ps.array[:] = srays
ph = rt.PinnedBuffer((n,), rt.HIT_DTYPE); po = rt.PinnedBuffer((n,), np.uint32)


def timed(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


t_page = timed(lambda: (s.trace_closest(rays, out=hits), s.trace_shadow(srays, out=occ)))
t_pin = timed(lambda: (s.trace_closest(pr.array, out=ph.array), s.trace_shadow(ps.array, out=po.array)))
assert hits.tobytes() == ph.array.tobytes()
print(json.dumps({"helpers": os.environ.get("B200RT_STAGING_HELPERS", "default"), "pageable_mrays": 2 * n / t_page / 1e6, "pinned_mrays": 2 * n / t_pin / 1e6}))

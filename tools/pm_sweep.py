"""Kernel-variant sweep of the photon-map lookups on one B200 (tuning aid): plain loop vs phased state machine, round length,
heaps in shared memory vs in the result array.  One process, CUDA events, results checked identical across variants.

    python tools/pm_sweep.py > gpurun_out/<tag>_pm_sweep.jsonl
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from libyafaray_b200 import pm, scenes

    pos, dirs = scenes.photon_cloud("surfaces", 1_000_000, seed=99)
    pts, nrm = scenes.gather_points(pos, 1_000_000, seed=98, jitter=0.002)
    m = pm.PhotonMap(pos, dirs)
    d_pts, d_nrm = torch.from_numpy(pts).cuda(), torch.from_numpy(nrm).cuda()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    reference = {}

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(reps):
            fn()
        ev[1].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]) / reps

    # (kernel, round_steps, smem_k, patience); kernel 0 = plain loop, 1 = phased, 2 = phased + one pop per step
    variants = [(0, 8, 0, 1), (0, 8, 256, 1), (1, 8, 0, 1), (1, 8, 0, 8), (2, 8, 0, 1), (2, 8, 0, 4), (2, 8, 0, 8), (2, 8, 0, 16), (2, 8, 0, 32), (2, 4, 0, 8), (2, 16, 0, 8),
                (2, 32, 0, 8), (2, 16, 0, 16), (2, 8, 256, 8)]
    cases = [("gather", 100, 2.5e-4), ("gather", 8, 1e-4), ("gather", 50, 1e-3), ("gather", 200, 1e-3)]
    for what, k, r2 in cases:
        for kernel, steps, smem_k, patience in variants:
            if smem_k and k > smem_k:
                continue
            pm.set_tuning(kernel, steps, smem_k, patience)
            if what == "gather":
                out = m.gather_device(d_pts, k, r2)
                ms = timed(lambda: m.gather_device(d_pts, k, r2, out=out))
                # signature on the device: counts, position-weighted photon ids and distance bits of the valid entries, radii bits
                cnt = out[1].long()
                weight = torch.arange(1, k + 1, device=d_pts.device)[None, :] * (torch.arange(k, device=d_pts.device)[None, :] < cnt[:, None])
                sig = (int(cnt.sum()), int((out[0][:, :, 0].long() * weight).sum()), int((out[0][:, :, 1].long() * weight).sum()), int(out[2].view(torch.int32).long().sum()))
            else:
                out = m.find_nearest_device(d_pts, d_nrm, r2)
                ms = timed(lambda: m.find_nearest_device(d_pts, d_nrm, r2, out=out))
                sig = (int(out.long().sum()),)
            key = (what, k, r2)
            same = reference.setdefault(key, sig) == sig
            print(json.dumps({"what": what, "k": k, "sq_radius": r2, "kernel": kernel, "round_steps": steps, "smem_k": smem_k, "patience": patience, "ms": ms,
                              "mpoints_per_s": len(pts) / ms / 1e3, "same_results_as_first_variant": same}), flush=True)
    m.close()


if __name__ == "__main__":
    main()

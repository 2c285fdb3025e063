"""Host-side sharding of a ray batch over the GPUs of one box.

The path has no exchange step: the scene is replicated, every rank traces a contiguous slice of the batch and the
slices are concatenated.  The only collectives are the optional all-gather of the result records and the MAX-reduce
of the elapsed time (multi-GPU numbers are the slowest rank's).  `trace_fn` is whatever traces one slice on the
rank's own device -- `rt.Scene.trace_closest` in production."""
from __future__ import annotations

import numpy as np


def shard_bounds(n: int, world: int):
    """Contiguous, balanced slices: rank r owns [bounds[r], bounds[r+1]).  Sizes differ by at most one ray."""
    if world < 1:
        raise ValueError("world must be >= 1")
    base, extra = divmod(int(n), world)
    sizes = [base + (1 if r < extra else 0) for r in range(world)]
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def shard_slice(n: int, rank: int, world: int) -> slice:
    b = shard_bounds(n, world)
    return slice(int(b[rank]), int(b[rank + 1]))


def trace_sharded(trace_fn, rays, rank: int, world: int, gather: bool = True, group=None):
    """Trace this rank's slice; with `gather`, return the whole batch's results on every rank (torch.distributed
    all_gather over whatever backend the process group uses: NCCL on GPUs, gloo in the CPU tests)."""
    sl = shard_slice(rays.shape[0], rank, world)
    mine = trace_fn(rays[sl])
    if world == 1 or not gather:
        return mine
    import torch
    import torch.distributed as dist
    bounds = shard_bounds(rays.shape[0], world)
    longest = int(np.max(np.diff(bounds)))
    raw = np.zeros(longest * mine.dtype.itemsize, np.uint8)
    raw[: mine.nbytes] = np.frombuffer(mine.tobytes(), np.uint8)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.from_numpy(raw).to(dev)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    out = []
    for r in range(world):
        n_r = int(bounds[r + 1] - bounds[r])
        out.append(np.frombuffer(parts[r].cpu().numpy().tobytes()[: n_r * mine.dtype.itemsize], dtype=mine.dtype))
    return np.concatenate(out)


def gather_sharded(gather_fn, points, k: int, rank: int, world: int, gather: bool = True, group=None):
    """The photon-map gather over the GPUs of one box: the photon map is replicated like the scene, every rank
    gathers a contiguous slice of the points, no exchange step.  gather_fn(points_slice) -> (found [m, k] (photon, dist_square),
    n_found [m], sq_radius_out [m]) -- `pm.PhotonMap.gather` with k and the radius bound, in production.  Returns the same three
    arrays for the whole batch (with `gather`) or for this rank's slice."""
    found_dtype = np.dtype([("photon", np.uint32), ("dist_square", np.float32)])
    row = np.dtype([("found", found_dtype, (k,)), ("n_found", np.uint32), ("radius", np.float32)])

    def packed(slice_points):
        found, n_found, radius = gather_fn(slice_points)
        out = np.zeros(len(slice_points), row)
        out["found"]["photon"], out["found"]["dist_square"] = found["photon"], found["dist_square"]
        out["n_found"], out["radius"] = n_found, radius
        return out

    whole = trace_sharded(packed, points, rank, world, gather=gather, group=group)
    return whole["found"], whole["n_found"], whole["radius"]


def max_over_ranks(value: float, world: int, device=None, group=None) -> float:
    if world == 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())

"""The N>1 host logic on CPU: world_size-2 gloo process groups.  The tracer plugged in here is the
oracle (test infrastructure); production plugs in rt.Scene.trace_closest on each rank's GPU."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from libyafaray_b200 import scenes, shard


def test_shard_bounds_are_balanced_and_cover():
    for n in (0, 1, 7, 1 << 20, (1 << 24) + 3):
        for world in (1, 2, 3, 4, 8):
            b = shard.shard_bounds(n, world)
            assert b[0] == 0 and b[-1] == n and len(b) == world + 1
            sizes = np.diff(b)
            assert sizes.min() >= 0 and sizes.max() - sizes.min() <= 1
    with pytest.raises(ValueError):
        shard.shard_bounds(10, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from oracle import kdo
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    xyz, idx, flags = scenes.heightfield(40)
    orc = kdo.Oracle(xyz, idx, flags)
    rays = scenes.rays_incoherent(10001, seed=3)  # odd size: ragged shards

    def trace(r):
        res = orc.trace_closest(r)
        out = np.zeros(r.shape[0], dtype=[("t", np.float32), ("u", np.float32), ("v", np.float32), ("prim", np.uint32)])
        out["t"], out["u"], out["v"] = res["t"], res["u"], res["v"]
        out["prim"] = res["prim"].astype(np.int64).astype(np.uint32)
        return out

    whole = shard.trace_sharded(trace, rays, rank, world, gather=True)
    slow = shard.max_over_ranks(float(rank + 1), world)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), whole)
    with open(os.path.join(out_dir, f"max{rank}.txt"), "w") as f:
        f.write(str(slow))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_trace_equals_unsharded(built, tmp_path):
    from oracle import kdo
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    xyz, idx, flags = scenes.heightfield(40)
    ref = kdo.Oracle(xyz, idx, flags).trace_closest(scenes.rays_incoherent(10001, seed=3))
    for rank in range(world):
        got = np.load(os.path.join(tmp_path, f"rank{rank}.npy"))
        assert got.shape[0] == 10001
        assert np.array_equal(got["t"], ref["t"]) and np.array_equal(got["prim"].astype(np.int64), ref["prim"].astype(np.int64) & 0xFFFFFFFF)
        assert float(open(os.path.join(tmp_path, f"max{rank}.txt")).read()) == float(world)


def _gather_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from oracle import pmo
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pos, dirs = scenes.photon_cloud("surfaces", 4000, seed=2)
    points, _ = scenes.gather_points(pos, 1501, seed=3)  # odd size: ragged shards
    o = pmo.OracleMap(pos, dirs)
    k, r2 = 12, 4e-3

    def gather(p):
        idx, d2, n_found, radius = o.gather(p, k, r2)
        found = np.zeros((len(p), k), dtype=[("photon", np.uint32), ("dist_square", np.float32)])
        found["photon"], found["dist_square"] = idx, d2
        return found, n_found, radius

    found, n_found, radius = shard.gather_sharded(gather, points, k, rank, world)
    np.savez(os.path.join(out_dir, f"gather{rank}.npz"), photon=found["photon"], d2=found["dist_square"], n_found=n_found, radius=radius)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_photon_gather_equals_unsharded(built, tmp_path):
    """Photon-map gather sharded by points over two ranks: the concatenated results are the unsharded ones."""
    from oracle import pmo
    world = 2
    mp.spawn(_gather_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    pos, dirs = scenes.photon_cloud("surfaces", 4000, seed=2)
    points, _ = scenes.gather_points(pos, 1501, seed=3)
    idx, d2, n_found, radius = pmo.OracleMap(pos, dirs).gather(points, 12, 4e-3)
    for rank in range(world):
        got = np.load(os.path.join(tmp_path, f"gather{rank}.npz"))
        assert np.array_equal(got["n_found"], n_found) and np.array_equal(got["radius"].view(np.uint32), radius.view(np.uint32))
        valid = np.arange(12)[None, :] < n_found[:, None]
        assert np.array_equal(got["photon"][valid], idx[valid]) and np.array_equal(got["d2"].view(np.uint32)[valid], d2.view(np.uint32)[valid])

#!/usr/bin/env python
"""Host<->device copy rates of this box (pinned memory, CUDA events), the ceiling of bench.py's `e2e` arm:
H2D alone, D2H alone, and both directions at once in the 1 GiB : 320 MiB proportion of one bench step."""
import json
import os
import torch

# under torchrun every rank measures ITS GPU at the same time (barrier before each timed loop, MAX time over ranks): the
# aggregate tells what the host's memory / PCIe root complexes feed to N GPUs at once -- the ceiling of bench.py's e2e at N > 1
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
GiB = 1 << 30
h_in = torch.empty(GiB, dtype=torch.uint8).pin_memory()
h_out = torch.empty(320 << 20, dtype=torch.uint8).pin_memory()
d_in = torch.empty(GiB, dtype=torch.uint8, device=dev)
d_out = torch.empty(320 << 20, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def h2d():
    with torch.cuda.stream(s1):
        s1.wait_stream(torch.cuda.current_stream())
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        s2.wait_stream(torch.cuda.current_stream())
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d(); d2h()


t_in, t_out, t_both = timed(h2d), timed(d2h), timed(both)
step_rays = 2 * (1 << 24)
if rank == 0:
    print(json.dumps({"n_gpus": world, "h2d_gbs": world * GiB / t_in / 1e9, "d2h_gbs": world * (320 << 20) / t_out / 1e9, "both_seconds_per_step": t_both,
                  "h2d_gbs_while_d2h": world * GiB / t_both / 1e9, "e2e_ceiling_mrays": world * step_rays / t_both / 1e6,
                  "numa": open("/sys/devices/system/node/online").read().strip() if os.path.exists("/sys/devices/system/node/online") else None, "cpus": os.cpu_count(),
                  "note": "one bench step moves 1 GiB of rays in and 320 MiB of results out; the ceiling assumes perfect overlap of both directions and zero kernel time"}))
if world > 1:
    dist.destroy_process_group()

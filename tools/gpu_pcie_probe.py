#!/usr/bin/env python
"""Raw host<->device copy bandwidth with every rank copying at the same time (run under torchrun): names the limiter of the
end-to-end numbers at N > 1 -- the host side of the PCIe fabric, not the transforms.  Each rank copies a pinned 1 GiB buffer
H2D, D2H and both ways at once (two streams), barrier-synchronised with the other ranks."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 28                                            # floats: 1 GiB
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_a = torch.empty(n, dtype=torch.float32, device="cuda")
d_b = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(kind, reps=4):
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if kind in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if kind in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gb = reps * n * 4 / 1e9 * (2 if kind == "both" else 1)
    t = torch.tensor([gb / dt], device="cuda", dtype=torch.float64)
    tot = t.clone()
    mn = t.clone()
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    return float(tot.item()), float(mn.item())


out = {"n_gpus": world, "bytes_per_copy": n * 4}
for kind in ("h2d", "d2h", "both"):
    run(kind, 1)
    tot, mn = run(kind)
    out[kind] = {"aggregate_gbs": tot, "slowest_rank_gbs": mn}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()

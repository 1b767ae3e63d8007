#!/usr/bin/env python3
"""How long does it take to give every rank one batch of frames (25 MB)? torchrun --nproc-per-node N tools/bcast_probe.py"""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
nbytes = 25267200
pad = (nbytes + world * 16 - 1) // (world * 16) * (world * 16)
buf = torch.zeros(pad, dtype=torch.uint8, device=dev)
part = torch.zeros(pad // world, dtype=torch.uint8, device=dev)


def timed(fn, n=20):
    for _ in range(5):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() * 1000.0


def bcast():
    dist.broadcast(buf, 0)


def scatter_gather():
    dist.scatter(part, list(buf.view(world, -1).unbind(0)) if rank == 0 else None, src=0)
    dist.all_gather_into_tensor(buf, part)


def allgather_only():
    dist.all_gather_into_tensor(buf, part)


r = {"world": world, "bytes": nbytes, "broadcast_us": timed(bcast), "scatter_allgather_us": timed(scatter_gather), "allgather_us": timed(allgather_only)}
small = torch.zeros(1024, dtype=torch.uint8, device=dev)
r["broadcast_1KB_us"] = timed(lambda: dist.broadcast(small, 0))
if rank == 0:
    print(r, flush=True)
dist.destroy_process_group()

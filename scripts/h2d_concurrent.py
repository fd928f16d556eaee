"""Aggregate pinned host -> device copy rate of this box with every rank copying at the same time (torchrun, one rank per GPU):
the wall the end-to-end leg of bench.py runs into at 8 GPUs.  1 GB pinned pieces, max-over-ranks time."""
import os
import time

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for mb, reps in ((256, 16), (1024, 8)):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if local == 0:
        gb = reps * (mb << 20) / 1e9
        print("H2D %d ranks x %4d MB pieces: %.1f GB/s per rank, %.1f GB/s aggregate" % (world, mb, gb / float(dt), world * gb / float(dt)), flush=True)
if world > 1:
    dist.destroy_process_group()

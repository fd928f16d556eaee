"""Pinned host -> device copy bandwidth of this box (what bounds the end-to-end leg of bench.py)."""
import time
import torch
for mb in (64, 256, 1024):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(8):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print("H2D %4d MB chunks: %.1f GB/s" % (mb, 8 * (mb << 20) / dt / 1e9))
    t = time.perf_counter()
    for _ in range(8):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print("D2H %4d MB chunks: %.1f GB/s" % (mb, 8 * (mb << 20) / dt / 1e9))

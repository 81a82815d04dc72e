"""Diagnostics: raw pinned H2D / D2H rates on this box, next to the e2e Detect call."""
import time
import torch
for mb in (1.9, 17.5, 256):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for direction in ("h2d", "d2h"):
        for _ in range(3):
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True)); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 20
        print(f"{direction} {mb:6.1f} MB: {dt * 1e3:.3f} ms  {n / dt / 1e9:.1f} GB/s")

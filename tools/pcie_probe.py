"""Diagnostics: raw pinned H2D / D2H rates on this box, with the default CPU placement and with the process bound to the GPU's
NUMA-local cores (NVML ideal affinity) before the pinned buffers are allocated."""
import os
import time
import torch


def probe(tag):
    for mb in (17.5, 256):
        n = int(mb * 1e6)
        h = torch.empty(n, dtype=torch.uint8).pin_memory()
        h.fill_(1)
        d = torch.empty(n, dtype=torch.uint8, device="cuda")
        for direction in ("h2d", "d2h"):
            for _ in range(3):
                (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True)); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
                torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 20
            print(f"{tag:18s} {direction} {mb:6.1f} MB: {dt * 1e3:.3f} ms  {n / dt / 1e9:.1f} GB/s")


torch.cuda.init()
print("affinity before:", len(os.sched_getaffinity(0)), "cpus", sorted(os.sched_getaffinity(0))[:4], "...")
probe("default placement")
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    n_words = (os.cpu_count() + 63) // 64
    mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
    cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
    print("NVML ideal affinity:", len(cpus), "cpus", sorted(cpus)[:4], "...")
    allowed = cpus & os.sched_getaffinity(0)
    if allowed:
        os.sched_setaffinity(0, allowed)
        probe("GPU-local cores")
    else:
        print("none of the GPU-local cpus is in this process's cpuset")
except Exception as e:          # noqa: BLE001
    print("nvml affinity unavailable:", repr(e))
try:
    print(open("/sys/devices/system/node/online").read().strip(), "NUMA nodes online")
except OSError:
    pass

"""Diagnostics: K2 reading conf straight from pinned host memory (zero-copy) vs cudaMemcpyAsync + K2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fdt_b200
from fdt_b200 import _lib, synth
B, C, CONF_T = 64, 2, 0.05
pri_np = synth.priors_numpy(640, 640)
loc_np, conf_np = synth.detect_inputs(B, pri_np, 20262, CONF_T, "random")
N = pri_np.shape[0]
dev = torch.device("cuda", 0)
L = _lib.lib()
conf_h = torch.from_numpy(conf_np).pin_memory()
conf_d = torch.empty_like(conf_h, device=dev)
ws = _lib.workspace(L.fdt_detect_workspace_bytes(B, N, C), dev, "diag")
st = _lib.stream_ptr()
cnt = torch.empty(B, dtype=torch.int32, device=dev)

def zc():
    _lib.check(L.fdt_detect_threshold_compact(conf_h.data_ptr(), B, N, C, CONF_T, ws.data_ptr(), ws.numel(), st))

def cp():
    conf_d.copy_(conf_h, non_blocking=True)
    _lib.check(L.fdt_detect_threshold_compact(conf_d.data_ptr(), B, N, C, CONF_T, ws.data_ptr(), ws.numel(), st))

for name, fn in (("copy + K2", cp), ("zero-copy K2", zc), ("copy + K2", cp), ("zero-copy K2", zc)):
    for _ in range(3):
        fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        fn(); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    _lib.check(L.fdt_detect_candidate_counts(ws.data_ptr(), ws.numel(), B, N, C, cnt.data_ptr(), st))
    print(f"{name:14s} {dt * 1e3:.3f} ms  ({conf_h.numel() * 4 / dt / 1e9:.1f} GB/s)  candidates {int(cnt.sum())}")

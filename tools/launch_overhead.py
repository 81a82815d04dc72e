"""Diagnostics: how much of k_sort_nms's event-measured time is launch overhead?  Times 1x and 2x back-to-back launches."""
import sys, os
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fdt_b200
from fdt_b200 import _lib, synth
import numpy as np

pri_np = synth.priors_numpy(640, 640)
loc_np, conf_np = synth.detect_inputs(64, pri_np, 20262, 0.05)
dev = torch.device("cuda")
loc, conf, pri = (torch.from_numpy(a).to(dev) for a in (loc_np, conf_np, pri_np))
B, N, C = 64, pri_np.shape[0], 2
L = _lib.lib()
out = torch.empty((B, C, 750, 5), device=dev)
ws = _lib.workspace(L.fdt_detect_workspace_bytes(B, N, C), dev, "x")
st = _lib.stream_ptr()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
s1 = lambda: _lib.check(L.fdt_detect_threshold_compact(conf.data_ptr(), B, N, C, 0.05, ws.data_ptr(), ws.numel(), st))
s2 = lambda: _lib.check(L.fdt_detect_sort_nms(loc.data_ptr(), pri.data_ptr(), B, N, C, 750, 5000, 0.3, 0.1, 0.2, out.data_ptr(), None, None, ws.data_ptr(), ws.numel(), st))

def run(n2, reps=30, flush_l2=True):
    ts = []
    for _ in range(reps):
        if flush_l2: flush.zero_()
        torch.cuda._sleep(400_000)
        s1()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n2): s2()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return np.median(ts)
for _ in range(3): s1(); s2()
a, b, c = run(1), run(2), run(4)
print(f"K3 x1 {a:.1f} us, x2 {b:.1f} us, x4 {c:.1f} us -> per launch {(c - a) / 3:.1f} us, fixed {a - (c - a) / 3:.1f} us")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda._sleep(400_000); e0.record(); e1.record(); torch.cuda.synchronize(); print("empty event pair", e0.elapsed_time(e1) * 1e3, "us")

def run_full(reps=50):
    ts = []
    for _ in range(reps):
        flush.zero_(); torch.cuda._sleep(400_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.fdt_detect(loc.data_ptr(), conf.data_ptr(), pri.data_ptr(), B, N, C, 750, 5000, 0.05, 0.3, 0.1, 0.2,
                                out.data_ptr(), None, None, ws.data_ptr(), ws.numel(), st))
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return np.median(ts)
print(f"fdt_detect (one call, no event between the stages): {run_full():.1f} us   [FDT_K3_PDL={os.environ.get('FDT_K3_PDL', 'default')}]")

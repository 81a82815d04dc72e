"""Diagnostics: where the event-timed Detect step goes (headline workload).  Each variant is timed like bench.py does
(L2 flush + spin outside the event pair, then the launches)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fdt_b200
from fdt_b200 import _lib, synth

B, C, TOP_K, NMS_TOP_K, CONF_T, NMS_T = 64, 2, 750, 5000, 0.05, 0.3
pri_np = synth.priors_numpy(640, 640)
loc_np, conf_np = synth.detect_inputs(B, pri_np, 20262, CONF_T, "random")
N = pri_np.shape[0]
dev = torch.device("cuda", 0)
loc, conf, pri = (torch.from_numpy(a).to(dev) for a in (loc_np, conf_np, pri_np))
L = _lib.lib()
out = torch.empty((B, C, TOP_K, 5), dtype=torch.float32, device=dev)
ws = _lib.workspace(L.fdt_detect_workspace_bytes(B, N, C), dev, "diag")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = _lib.stream_ptr()


def k2():
    _lib.check(L.fdt_detect_threshold_compact(conf.data_ptr(), B, N, C, CONF_T, ws.data_ptr(), ws.numel(), st))


def k3():
    _lib.check(L.fdt_detect_sort_nms(loc.data_ptr(), pri.data_ptr(), B, N, C, TOP_K, NMS_TOP_K, NMS_T, 0.1, 0.2,
                                     out.data_ptr(), None, None, ws.data_ptr(), ws.numel(), st))


def full():
    _lib.check(L.fdt_detect(loc.data_ptr(), conf.data_ptr(), pri.data_ptr(), B, N, C, TOP_K, NMS_TOP_K, CONF_T, NMS_T, 0.1, 0.2,
                            out.data_ptr(), None, None, ws.data_ptr(), ws.numel(), st))


def timed(fn, do_flush=True, reps=40, pre=None):
    ts = []
    for i in range(reps + 5):
        if do_flush:
            flush.zero_()
        torch.cuda._sleep(250_000)
        if pre:
            pre()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if i >= 5:
            ts.append(a.elapsed_time(b) * 1e3)
    return "%.1f us (min %.1f)" % (np.mean(ts), np.min(ts))


small = torch.empty(64, dtype=torch.int32, device=dev)
print("empty event pair          ", timed(lambda: None))
print("memset 768 B              ", timed(lambda: small.zero_()))
print("K2 (memset + kernel) cold ", timed(k2))
print("K2 warm L2                ", timed(k2, do_flush=False))
print("K3 alone cold (after K2)  ", timed(k3, pre=k2))
print("K3 alone warm             ", timed(k3, do_flush=False, pre=k2))
print("full step cold            ", timed(full))
print("full step warm            ", timed(full, do_flush=False))

# the same step as a CUDA graph (stream capture keeps the programmatic-dependency edges)
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    st_side = _lib.stream_ptr()

    def full_side():
        _lib.check(L.fdt_detect(loc.data_ptr(), conf.data_ptr(), pri.data_ptr(), B, N, C, TOP_K, NMS_TOP_K, CONF_T, NMS_T, 0.1, 0.2,
                                out.data_ptr(), None, None, ws.data_ptr(), ws.numel(), st_side))
    full_side()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=side):
        full_side()
ref = out.clone()
out.zero_()
g.replay()
torch.cuda.synchronize()
print("graph replay identical:", bool(torch.equal(ref, out)))
print("full step as graph cold   ", timed(g.replay))
print("full step as graph warm   ", timed(g.replay, do_flush=False))

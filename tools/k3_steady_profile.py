"""Diagnostics: average per-phase cycles of k_sort_nms over EVERY CTA of many overlapping Detect calls (FDT_K3_PROFILE=2): where
the SM time of the steady state goes.  python tools/k3_steady_profile.py [calls] [depth]"""
import ctypes as C
import os
import sys
os.environ["FDT_K3_PROFILE"] = "2"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fdt_b200
from fdt_b200 import _lib, synth

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 60
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pri = synth.priors_numpy(640, 640)
B, N = 64, pri.shape[0]
sets = [tuple(torch.from_numpy(a).cuda() for a in synth.detect_inputs(B, pri, 20262 + 1000 * r, 0.05)) for r in range(3)]
p = torch.from_numpy(pri).cuda()
L = _lib.lib()
ws = torch.empty(L.fdt_detect_workspace_bytes_depth(B, N, 2, depth), dtype=torch.uint8, device="cuda")
outs = [torch.empty((B, 2, 750, 5), device="cuda") for _ in range(4)]
st = _lib.stream_ptr()


def run(n):
    torch.cuda._sleep(3_000_000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        l, c = sets[i % 3]
        _lib.check(L.fdt_detect(l.data_ptr(), c.data_ptr(), p.data_ptr(), B, N, 2, 750, 5000, 0.05, 0.3, 0.1, 0.2, outs[i % 4].data_ptr(), None, None,
                                ws.data_ptr(), ws.numel(), st))
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


run(5)
out = (C.c_longlong * 1024)()
_lib.check(L.fdt_debug_k3_profile(out))          # reads and clears
ms = run(calls)
_lib.check(L.fdt_debug_k3_profile(out))
n = max(out[30], 1)
names = {0: "range / row scan setup", 1: "hist + scan (+ select)  [fused: first row scan]", 2: "scatter (+ bitonic)", 5: "win: rank", 6: "win: decode + geometry",
         7: "win: csr build", 8: "win: A kept-query", 9: "win: B window-query", 10: "win: resolve", 16: "win: append", 11: "output"}
print(f"{calls} calls, depth {depth}: {ms * 1e3:.2f} us per call with profiling barriers; {n} CTAs profiled")
tot = 0
for i, nm in names.items():
    v = out[i] / n
    tot += v
    print(f"  {nm:50s} {v:9.0f} cyc  {v / 1965:6.2f} us")
print(f"  sum {tot:.0f} cyc = {tot / 1965:.2f} us per CTA; stage1..append measured directly: {out[31] / n / 1965:.2f} us")
print(f"  kept {out[12] / n:.1f}  k {out[13] / n:.1f}  rounds {out[14] / n:.2f}  sweeps {out[15] / n:.2f}")
print(f"  SM time per call = 64 CTAs x {tot / 1965:.2f} us = {64 * tot / 1965:.0f} SM-us -> {64 * tot / 1965 / 148:.2f} us per call if the 148 SMs were always busy")

"""Diagnostics: per-phase cycle breakdown of k_sort_nms (CTA 0) on the headline workload.  FDT_K3_PROFILE=1 python tools/k3_profile.py [mode] [batch]"""
import ctypes as C
import os
import sys
os.environ["FDT_K3_PROFILE"] = "1"
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fdt_b200
from fdt_b200 import _lib, synth
from fdt_b200.layers import Detect

mode = sys.argv[1] if len(sys.argv) > 1 else "random"
pri = synth.priors_numpy(640, 640)
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
loc, conf = synth.detect_inputs(B, pri, 20262, 0.05, mode)
det = Detect(2, 0, 750, 0.05, 0.3)
args = [torch.from_numpy(a).cuda() for a in (loc, conf, pri)]
for _ in range(3):
    det(*args)
torch.cuda.synchronize()
out = (C.c_longlong * 1024)()
_lib.check(_lib.lib().fdt_debug_k3_profile(out))
names = ["minmax", "hist+scan(+select)", "scatter(+bitonic)", "-", "-", "win:rank", "win:decode+geom", "win:csr build",
         "win:A kept-query", "win:B window-query", "win:resolve", "output", "kept", "k", "rounds", "sweeps", "win:append",
         "look-calls", "sum nd", "ovf threads", "bigcell cands", "t(sweep1)", "t(sweep2)", "t(sweep3)", "t(sweep4)", "max visits", "sum visits", "max tests", "sum tests"]
tot = sum(out[i] for i in list(range(12)) + [16])
for i, n in enumerate(names):
    if n in ("rounds", "kept", "k", "-", "sweeps", "look-calls", "sum nd", "ovf threads", "bigcell cands", "max visits", "sum visits", "max tests", "sum tests") or n.startswith("t("):
        print(f"{n:14s} {out[i]}")
    else:
        print(f"{n:14s} {out[i]:9d} cyc  {100 * out[i] / max(tot, 1):5.1f}%")
print("total", tot, "cycles =", tot / 1.965e3, "us @1.965GHz")

import numpy as np
cyc = np.array([out[64 + i] for i in range(256)]); cyc = cyc[cyc > 0]
meta = np.array([out[320 + i] for i in range(256)])[:len(cyc)]
print("per-CTA cycles: n=%d min=%d mean=%d max=%d (%.1f us)  rounds max=%d  k min/max=%d/%d" % (len(cyc), cyc.min(), cyc.mean(), cyc.max(), cyc.max() / 1965.0, (meta // 100000).max(), (meta % 100000).min(), (meta % 100000).max()))
print("K2 on the same clock (ns): first block start %+d, last block end %+d" % (out[44] - out[40], out[45] - out[40]))
print("globaltimer (ns): first CTA start 0, last CTA start +%d, first CTA end +%d, last CTA end +%d" % (out[41] - out[40], out[42] - out[40], out[43] - out[40]))

print("phase B per warp (CTA 0): cycles", [out[640 + i] for i in range(32)])
print("   visits sum", [out[672 + i] for i in range(32)])
print("   visits max-lane", [out[704 + i] for i in range(32)])
print("   chunks", [out[736 + i] for i in range(32)])

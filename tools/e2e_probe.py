"""Diagnostics: end-to-end Detect call on pinned host tensors (fdt_detect_host), ms per call; FDT_HOST_OUT_COPY=1 stages the output."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fdt_b200
from fdt_b200 import synth
from fdt_b200.layers import Detect
pri_np = synth.priors_numpy(640, 640)
loc_np, conf_np = synth.detect_inputs(64, pri_np, 20262, 0.05, "random")
loc_h, conf_h, pri_h = (torch.from_numpy(a).pin_memory() for a in (loc_np, conf_np, pri_np))
det = Detect(2, 0, 750, 0.05, 0.3)
for _ in range(5):
    det(loc_h, conf_h, pri_h)
per = []
for _ in range(50):
    t0 = time.perf_counter(); det(loc_h, conf_h, pri_h); per.append(time.perf_counter() - t0)
print("FDT_HOST_OUT_COPY=%s: median %.3f ms, mean %.3f ms, min %.3f ms" % (os.environ.get("FDT_HOST_OUT_COPY", "0"), 1e3 * np.median(per), 1e3 * np.mean(per), 1e3 * np.min(per)))
d = torch.empty(conf_h.numel() * 4, dtype=torch.uint8, device="cuda")
src = conf_h.view(torch.uint8).view(-1)
for _ in range(3):
    d.copy_(src, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    d.copy_(src, non_blocking=True); torch.cuda.synchronize()
print("raw H2D of conf (17.5 MB): %.3f ms" % ((time.perf_counter() - t0) / 20 * 1e3))

"""Diagnostics: Detect steps issued back to back over rotating input sets larger than L2 (no flush, one event pair around all steps)
vs the per-step flushed timing of bench.py."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fdt_b200
from fdt_b200 import _lib, synth

B, C, TOP_K, NMS_TOP_K, CONF_T, NMS_T = 64, 2, 750, 5000, 0.05, 0.3
R = int(sys.argv[1]) if len(sys.argv) > 1 else 3
K = int(sys.argv[2]) if len(sys.argv) > 2 else 100
pri_np = synth.priors_numpy(640, 640)
N = pri_np.shape[0]
dev = torch.device("cuda", 0)
sets = []
for r in range(R):
    loc_np, conf_np = synth.detect_inputs(B, pri_np, 20262 + 1000 * r, CONF_T, "random")
    sets.append((torch.from_numpy(loc_np).to(dev), torch.from_numpy(conf_np).to(dev)))
pri = torch.from_numpy(pri_np).to(dev)
L = _lib.lib()
out = torch.empty((B, C, TOP_K, 5), dtype=torch.float32, device=dev)
ws = _lib.workspace(L.fdt_detect_workspace_bytes(B, N, C), dev, "diag")
st = _lib.stream_ptr()


def step(i):
    loc, conf = sets[i % R]
    _lib.check(L.fdt_detect(loc.data_ptr(), conf.data_ptr(), pri.data_ptr(), B, N, C, TOP_K, NMS_TOP_K, CONF_T, NMS_T, 0.1, 0.2,
                            out.data_ptr(), None, None, ws.data_ptr(), ws.numel(), st))


for i in range(10):
    step(i)
torch.cuda.synchronize()
for trial in range(3):
    torch.cuda._sleep(2_000_000)                 # ~1 ms head start for the host
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(K):
        step(i)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / K
    print(f"{R} rotating sets ({R * 54.9:.0f} MB), {K} steps back to back: {ms * 1e3:.1f} us/step -> {B / ms * 1e3 / 1e6:.3f} M frames/s")

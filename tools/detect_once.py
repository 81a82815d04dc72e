"""Runs the headline Detect a few times (target for `ncu -k regex:k_sort_nms`).  python tools/detect_once.py [mode] [batch] [reps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fdt_b200
from fdt_b200 import synth
from fdt_b200.layers import Detect

mode = sys.argv[1] if len(sys.argv) > 1 else "random"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
pri = synth.priors_numpy(640, 640)
loc, conf = synth.detect_inputs(B, pri, 20262, 0.05, mode)
det = Detect(2, 0, 750, 0.05, 0.3)
args = [torch.from_numpy(a).cuda() for a in (loc, conf, pri)]
for _ in range(reps):
    out = det(*args)
torch.cuda.synchronize()
print("ok", tuple(out.shape), float(out.sum()))

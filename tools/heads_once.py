"""Runs Detect.detect_heads on the headline shape a few times (target for ncu).  python tools/heads_once.py [batch] [reps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fdt_b200
from fdt_b200 import synth
from fdt_b200.layers import Detect, heads_to_loc_conf

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
loc_maps, conf_maps, neg_max = synth.head_maps(B, 640, 640, 6060)
pri = torch.from_numpy(synth.priors_numpy(640, 640)).cuda()
lm, cm = [torch.from_numpy(m).cuda() for m in loc_maps], [torch.from_numpy(m).cuda() for m in conf_maps]
det = Detect(2, 0, 750, 0.05, 0.3)
for _ in range(reps):
    out = det.detect_heads(lm, cm, pri)
    out2 = det(*heads_to_loc_conf(lm, cm, neg_max), pri)
torch.cuda.synchronize()
print("ok", tuple(out.shape), bool(torch.equal(out, out2)))

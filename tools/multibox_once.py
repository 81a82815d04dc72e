"""Runs the config-3 MultiBoxLoss forward a few times (target for ncu).  python tools/multibox_once.py [reps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fdt_b200
from fdt_b200 import synth
from fdt_b200.layers import MultiBoxLoss

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
pri = synth.priors_numpy(640, 640)
loc, conf, targets = synth.multibox_inputs(32, pri, 3030, 0, 200)
l, c, p = (torch.from_numpy(a).cuda() for a in (loc, conf, pri))
tg = [torch.from_numpy(t).cuda() for t in targets]
crit = MultiBoxLoss(2, 0.35, True, 0, True, 3, 0.35, False, bipartite=False)
for _ in range(reps):
    ll, lc = crit((l, c, p), tg)
torch.cuda.synchronize()
print("ok", float(ll), float(lc))

"""Diagnostics: MultiBoxLoss GPU vs oracle on one synthetic batch, piece by piece.  python tools/mbl_diff.py B seed g_lo g_hi"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fdt_b200
from fdt_b200 import synth
from fdt_b200.layers import MultiBoxLoss
from oracle import oracle as orc
B, seed, g_lo, g_hi = (int(a) for a in sys.argv[1:5])
pri = synth.priors_numpy(640, 640)
loc, conf, targets = synth.multibox_inputs(B, pri, seed, g_lo, g_hi)
print("G per image", [t.shape[0] for t in targets])
cu = lambda a: torch.from_numpy(a).cuda()
crit = MultiBoxLoss(2, 0.35, True, 0, True, 3, 0.35, False, bipartite=False)
ll, lc = crit((cu(loc), cu(conf), cu(pri)), [cu(t) for t in targets])
loc_t, conf_t, sel = (t.cpu().numpy() for t in crit.last_aux)
r = orc.multibox_loss(loc, conf, pri, targets, 0.35, 3, False)
print("gpu", float(ll), float(lc), "oracle", r["loss_l"], r["loss_c"])
d = conf_t != r["conf_t"]
print("conf_t mismatches", int(d.sum()), "per image", d.sum(1).tolist())
for b in np.where(d.any(1))[0][:3]:
    idx = np.where(d[b])[0][:8]
    print(" image", b, "G", targets[b].shape[0], "priors", idx.tolist(), "gpu", conf_t[b, idx].tolist(), "oracle", r["conf_t"][b, idx].tolist())
    print("  targets head", targets[b][:3].tolist())
selr = r["neg"] | (r["conf_t"] > 0)
print("sel mismatches", int((sel.astype(bool) != selr).sum()))
pos = (conf_t > 0) & (r["conf_t"] > 0)
print("loc_t max abs diff on common positives", float(np.abs(loc_t[pos] - r["loc_t"][pos]).max()) if pos.any() else None)

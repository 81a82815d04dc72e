import sys; import torch, numpy as np
import fdt_b200
from fdt_b200 import synth, tracker as T
frames = synth.tracker_frames(F=1500, seed=4040, d_lo=1, d_hi=300, n_objects=300, empty_every=0)
dets, off = T.pack_frames(frames)
for _ in range(2):
    r = T.iou_track_raw(dets, off)
torch.cuda.synchronize()
print(len(r[2]))

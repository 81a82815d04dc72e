"""What the reference's scripts do with Detect's output (My_test.py:43-72, iouTracke_cal.py:55-84), on the GPU:
`detections_to_rows` = per image, class by class, the leading rows with score >= threshold, boxes scaled to pixels."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib


def detections_to_rows(detections, threshold, width, height, dummy_if_empty=False):
    """detections [B, C, top_k, 5] (Detect output) -> list of B float32 arrays [n_b, 5] rows [x1, y1, x2, y2, score] in pixels.
    The python loop of the reference stops at the first row whose score is below `threshold`; so does this.
    dummy_if_empty=True returns the reference's float64 placeholder [[0, 0, 0, 0, 0.4]] for an image without detections
    (My_test.py:62-63, iouTracke_cal.py:73-74)."""
    dev = _lib.require_cuda()
    d = _lib.dev_f32(detections, detections.device if detections.is_cuda else dev)
    B, C_, K, _ = d.shape
    rows = torch.empty((B, C_ * K, 5), dtype=torch.float32, device=d.device)
    n = torch.empty(B, dtype=torch.int32, device=d.device)
    with torch.cuda.device(d.device):
        _lib.check(_lib.lib().fdt_detections_to_rows(_lib.ptr(d), B, C_, K, float(threshold), float(width), float(height),
                                                     _lib.ptr(rows), _lib.ptr(n), _lib.stream_ptr()))
    n_h = n.cpu().numpy()
    rows_h = rows.cpu().numpy()
    out = []
    for b in range(B):
        if n_h[b] == 0 and dummy_if_empty:
            out.append(np.array([[0, 0, 0, 0, 0.4]]))
        else:
            out.append(rows_h[b, :n_h[b]].copy())
    return out

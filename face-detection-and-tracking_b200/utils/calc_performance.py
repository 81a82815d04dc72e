"""Drop-in for the tracker-side functions of the reference's utils/calc_performance.py (float64 numpy in,
float64 numpy out), computed by fdt_calculate_iou_f64 on the GPU."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib


def calculate_iou(box_a, box_b):
    """Pairwise jaccard overlap of [A,4] and [B,4] corner boxes -> ndarray [A,B]   [calc_performance.py:54-74]"""
    dev = _lib.require_cuda()
    a = torch.as_tensor(np.ascontiguousarray(box_a, dtype=np.float64)).to(dev)
    b = torch.as_tensor(np.ascontiguousarray(box_b, dtype=np.float64)).to(dev)
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().fdt_calculate_iou_f64(_lib.ptr(a), a.shape[0], _lib.ptr(b), b.shape[0], _lib.ptr(out),
                                                    _lib.stream_ptr()))
    return out.cpu().numpy()

"""Shared launcher for the sibling NMS variants (fdt_nms_variant).  No CPU fallback."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib


def nms_variant(boxes, scores, thresh: float, flags: int, keep_dtype: bool = False) -> torch.Tensor:
    """boxes[n,4], scores[n] (numpy or torch, any device) -> kept indices, int64 CUDA tensor in keep order.
    fp32 arithmetic (float64 inputs are rounded to fp32 first) unless keep_dtype=True and the inputs are float64: then the
    overlap rule is evaluated in float64 (fdt_nms_variant_f64), as numpy does for MTCNN's float64 `dets`."""
    dev = _lib.require_cuda()
    b = torch.as_tensor(np.ascontiguousarray(boxes) if isinstance(boxes, np.ndarray) else boxes)
    s = torch.as_tensor(np.ascontiguousarray(scores) if isinstance(scores, np.ndarray) else scores)
    if keep_dtype and b.dtype == torch.float64:
        b = b.reshape(-1, 4).to(dev).contiguous()
        s = s.reshape(-1).to(device=dev, dtype=torch.float64).contiguous()
        n = s.shape[0]
        if b.shape[0] != n:
            raise ValueError(f"nms: {b.shape[0]} boxes for {n} scores")
        keep = torch.zeros(n, dtype=torch.int64, device=dev)
        if n == 0:
            return keep
        count = torch.zeros(1, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            L = _lib.lib()
            ws = _lib.workspace(L.fdt_nms_f64_workspace_bytes(n), dev, "nms64")
            _lib.check(L.fdt_nms_variant_f64(_lib.ptr(b), _lib.ptr(s), n, float(thresh), int(flags), _lib.ptr(keep), _lib.ptr(count),
                                             _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
        return keep[:int(count.item())]
    b = _lib.dev_f32(b.reshape(-1, 4), dev)
    s = _lib.dev_f32(s.reshape(-1), dev)
    n = s.shape[0]
    if b.shape[0] != n:
        raise ValueError(f"nms: {b.shape[0]} boxes for {n} scores")
    keep = torch.zeros(n, dtype=torch.int64, device=dev)
    if n == 0:
        return keep
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        L = _lib.lib()
        ws = _lib.workspace(L.fdt_nms_workspace_bytes(n), dev, "nms")
        _lib.check(L.fdt_nms_variant(_lib.ptr(b), _lib.ptr(s), n, float(thresh), int(flags), _lib.ptr(keep), _lib.ptr(count),
                                     _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    return keep[:int(count.item())]

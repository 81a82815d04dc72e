"""PriorBoxLayer -- drop-in for layers/functions/prior_box.py:9-44, computed by fdt_priorbox on the GPU."""
from __future__ import annotations

import ctypes as C
from math import sqrt

import torch

from ... import _lib


class PriorBoxLayer:
    """Centre-form priors [cx, cy, w, h] of one pyramid level, y outer / x inner (prior_box.py:31-42).

    Same constructor and call signature as the reference.  The result lives on the current CUDA
    device (the reference's models move it there right after building it, pyramid.py:283)."""

    def __init__(self, width, height, stride=(4, 8, 16, 32, 64, 128), box=(16, 32, 64, 128, 256, 512),
                 scale=(1, 1, 1, 1, 1, 1), aspect_ratios=([], [], [], [], [], [])):
        self.width = width
        self.height = height
        self.stride = stride
        self.box = box
        self.scales = scale
        self.aspect_ratios = aspect_ratios

    def __call__(self, prior_idx, f_width, f_height):
        dev = _lib.require_cuda()
        n_scales = int(self.scales[prior_idx])
        ars = list(self.aspect_ratios[prior_idx])
        # python-float arithmetic exactly as prior_box.py:33 and :41, handed to the kernel as fp64
        box_scale = (C.c_double * max(n_scales, 1))(*[(2 ** (1 / 3)) ** s for s in range(n_scales)])
        sqrt_ar = (C.c_double * max(len(ars), 1))(*[sqrt(ar) for ar in ars])
        n = int(f_height) * int(f_width) * n_scales * (1 + len(ars))
        out = torch.empty((n, 4), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().fdt_priorbox(float(self.width), float(self.height), float(self.stride[prior_idx]),
                                           float(self.box[prior_idx]), n_scales, C.addressof(box_scale), len(ars),
                                           C.addressof(sqrt_ar), int(f_width), int(f_height), _lib.ptr(out),
                                           _lib.stream_ptr()))
        return out

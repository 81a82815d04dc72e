"""Head post-processing that feeds Detect / MultiBoxLoss.

The reference does this inline in every model's forward (pyramid.py:291-309 and :331-332, same code in
pyramid_mobile_try1.py:297-327 and pyramid_mb2_try3/4/5.py): per pyramid level the 4-channel confidence map is
reduced to (neg, pos) by max-in-out, both maps are permuted NCHW -> NHWC, flattened and concatenated over the
levels into loc[B,N,4] / conf[B,N,2], and the test phase applies a 2-way softmax before Detect.

`heads_to_loc_conf` materialises exactly those tensors with one kernel; `Detect.detect_heads` (detection.py here)
skips the materialisation altogether.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from ... import _lib


def _level_args(loc_maps, conf_maps, neg_max):
    """-> (device, B, N, keep-alive tensors, ctypes arrays) for the C ABI's per-level description."""
    _lib.require_cuda()
    L = len(conf_maps)
    if L == 0 or (loc_maps is not None and len(loc_maps) != L):
        raise ValueError("heads: need one loc map and one conf map per pyramid level")
    dev = conf_maps[0].device
    if not conf_maps[0].is_cuda:
        raise ValueError("heads: the maps must be CUDA tensors (there is no CPU path)")
    B = conf_maps[0].size(0)
    conf = [_lib.dev_f32(m, dev) for m in conf_maps]
    loc = [_lib.dev_f32(m, dev) for m in loc_maps] if loc_maps is not None else None
    for l, m in enumerate(conf):
        if m.dim() != 4 or m.size(1) != 4 or m.size(0) != B:
            raise ValueError(f"heads: conf map {l} must be [B,4,H,W], got {tuple(m.shape)}")
        if loc is not None and tuple(loc[l].shape) != tuple(m.shape):
            raise ValueError(f"heads: loc map {l} must match the conf map {tuple(m.shape)}, got {tuple(loc[l].shape)}")
    if neg_max is None:
        neg_max = [1] + [0] * (L - 1)                    # level 0: three negatives, the others three positives
    fh = (C.c_int * L)(*[int(m.size(2)) for m in conf])
    fw = (C.c_int * L)(*[int(m.size(3)) for m in conf])
    nm = (C.c_int * L)(*[int(bool(v)) for v in neg_max])
    cp = (C.c_void_p * L)(*[m.data_ptr() for m in conf])
    lp = (C.c_void_p * L)(*[m.data_ptr() for m in loc]) if loc is not None else None
    N = sum(int(m.size(2)) * int(m.size(3)) for m in conf)
    return dev, B, N, (conf, loc), (lp, cp, fh, fw, nm, L)


def heads_to_loc_conf(loc_maps, conf_maps, neg_max=None, softmax=True):
    """Per-level NCHW maps -> (loc[B,N,4], conf[B,N,2]) as pyramid.py:291-309 builds them; softmax=True also
    applies nn.Softmax(dim=-1) (:332, test phase), softmax=False keeps the raw logits MultiBoxLoss consumes."""
    dev, B, N, keep, (lp, cp, fh, fw, nm, L) = _level_args(loc_maps, conf_maps, neg_max)
    with torch.cuda.device(dev):
        loc = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
        conf = torch.empty((B, N, 2), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().fdt_heads_to_loc_conf(lp, cp, fh, fw, nm, L, B, 1 if softmax else 0,
                                                    _lib.ptr(loc), _lib.ptr(conf), _lib.stream_ptr()))
    del keep
    return loc, conf

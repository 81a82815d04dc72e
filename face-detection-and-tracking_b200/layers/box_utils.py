"""Drop-in for the reference's layers/box_utils.py, every function backed by a libfdt_b200 kernel.

Same names, argument order and in-place output conventions as the reference.  Inputs may be CPU or
CUDA tensors; compute always happens on the current CUDA device and results come back on the
device of the first tensor argument.  There is no CPU fallback.
"""
from __future__ import annotations

import torch

from .. import _lib
from ..data import face  # noqa: F401  (box_utils.py:3 imports it too)


def _prep(*tensors):
    dev = _lib.require_cuda()
    src = tensors[0].device
    if src.type == "cuda":
        dev = src
    return dev, src, [_lib.dev_f32(t, dev) for t in tensors]


def _elementwise(fn_name, boxes, *extra_tensors, scalars=()):
    dev, src, ts = _prep(boxes, *extra_tensors)
    out = torch.empty_like(ts[0])
    with torch.cuda.device(dev):
        fn = getattr(_lib.lib(), fn_name)
        _lib.check(fn(*[_lib.ptr(t) for t in ts], ts[0].shape[0], *scalars, _lib.ptr(out), _lib.stream_ptr()))
    return out.to(src)


def point_form(boxes):
    """(cx, cy, w, h) -> (xmin, ymin, xmax, ymax)   [box_utils.py:7-16]"""
    return _elementwise("fdt_point_form", boxes)


def center_size(boxes):
    """(xmin, ymin, xmax, ymax) -> (cx, cy, w, h)   [box_utils.py:19-28]"""
    return _elementwise("fdt_center_size", boxes)


def _pairwise(fn_name, box_a, box_b):
    dev, src, (a, b) = _prep(box_a, box_b)
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        fn = getattr(_lib.lib(), fn_name)
        _lib.check(fn(_lib.ptr(a), a.shape[0], _lib.ptr(b), b.shape[0], _lib.ptr(out), _lib.stream_ptr()))
    return out.to(src)


def intersect(box_a, box_b):
    """Pairwise intersection area [A,B]   [box_utils.py:31-67, the regular branch; the reference's
    >1000 MB CPU branch (:44-56) is buggy and outside every configured size]"""
    return _pairwise("fdt_intersect", box_a, box_b)


def calculate_iou(box_a, box_b):
    """Pairwise jaccard overlap [A,B] = inter / ((area_a + area_b) - inter)   [box_utils.py:70-100]"""
    return _pairwise("fdt_calculate_iou", box_a, box_b)


def encode(matched, priors, variances):
    """[box_utils.py:213-234]"""
    return _elementwise("fdt_encode", matched, priors, scalars=(float(variances[0]), float(variances[1])))


def decode(loc, priors, variances):
    """[box_utils.py:238-258]"""
    return _elementwise("fdt_decode", loc, priors, scalars=(float(variances[0]), float(variances[1])))


def log_sum_exp(x):
    """log(sum(exp(x - x.max()), 1, keepdim=True)) + x.max()   [box_utils.py:261-269]"""
    dev, src, (xx,) = _prep(x)
    out = torch.empty((xx.shape[0], 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ws = _lib.workspace(256, dev, "lse")
        _lib.check(_lib.lib().fdt_log_sum_exp(_lib.ptr(xx), xx.shape[0], xx.shape[1], _lib.ptr(out), _lib.ptr(ws),
                                              ws.numel(), _lib.stream_ptr()))
    return out.to(src)


def nms(boxes, scores, overlap=0.5, top_k=200):
    """Greedy NMS   [box_utils.py:275-340].  -> (keep: int64[n] zero-padded, count: int)
    The reference returns count as a python int, so this call synchronises on it."""
    dev, src, (b, s) = _prep(boxes.reshape(-1, 4), scores.reshape(-1))
    n = s.shape[0]
    keep = torch.zeros(n, dtype=torch.int64, device=dev)
    if b.numel() == 0:
        return keep.to(src), 0
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        L = _lib.lib()
        ws = _lib.workspace(L.fdt_nms_workspace_bytes(n), dev, "nms")
        _lib.check(L.fdt_nms(_lib.ptr(b), _lib.ptr(s), n, float(overlap), int(top_k), _lib.ptr(keep), _lib.ptr(count),
                             _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    return keep.to(src), int(count.item())


def _match(bipartite, threshold, truth_loc, priors, variances, truth_conf, loc_t, conf_t, idx):
    if truth_loc.shape[0] == 0:
        raise IndexError("max(): Expected reduction dim 0 to have non-zero size.")   # what the reference does (Q3)
    dev, _, (t, p, lab) = _prep(truth_loc.reshape(-1, 4), priors, truth_conf.reshape(-1))
    N = p.shape[0]
    gt = torch.cat([t, lab.unsqueeze(1)], 1).contiguous()
    off = torch.tensor([0, t.shape[0]], dtype=torch.int64, device=dev)
    lt = torch.empty((1, N, 4), dtype=torch.float32, device=dev)
    ct = torch.empty((1, N), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        L = _lib.lib()
        ws = _lib.workspace(L.fdt_match_workspace_bytes(1, N, t.shape[0]), dev, "match")
        _lib.check(L.fdt_match_encode(_lib.ptr(p), _lib.ptr(gt), _lib.ptr(off), t.shape[0], 1, N, float(threshold),
                                      float(variances[0]), float(variances[1]), int(bool(bipartite)),
                                      _lib.ptr(lt), _lib.ptr(ct), None, None, _lib.ptr(ws), ws.numel(),
                                      _lib.stream_ptr()))
    loc_t[idx] = lt[0].to(loc_t.device)          # box_utils.py:161-162 / 209-210 in-place convention
    conf_t[idx] = ct[0].to(conf_t.device)


def match_ensure_max_prior(threshold, truth_loc, priors, variances, truth_conf, loc_t, conf_t, idx):
    """[box_utils.py:103-162]"""
    _match(True, threshold, truth_loc, priors, variances, truth_conf, loc_t, conf_t, idx)


def match_default(threshold, truth_loc, priors, variances, truth_conf, loc_t, conf_t, idx):
    """[box_utils.py:165-210]"""
    _match(False, threshold, truth_loc, priors, variances, truth_conf, loc_t, conf_t, idx)

"""Drop-in for the tracker-side functions of the reference's utils/calc_performance.py (float64 numpy in,
float64 numpy out), computed by fdt_calculate_iou_f64 on the GPU."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib


def _pairwise(fn_name, box_a, box_b):
    dev = _lib.require_cuda()
    a = torch.as_tensor(np.ascontiguousarray(box_a, dtype=np.float64)).to(dev)
    b = torch.as_tensor(np.ascontiguousarray(box_b, dtype=np.float64)).to(dev)
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(getattr(_lib.lib(), fn_name)(_lib.ptr(a), a.shape[0], _lib.ptr(b), b.shape[0], _lib.ptr(out), _lib.stream_ptr()))
    return out.cpu().numpy()


def intersect(box_a, box_b):
    """Pairwise intersection area of [A,4] and [B,4] corner boxes -> ndarray [A,B]   [calc_performance.py:4-31]"""
    return _pairwise("fdt_intersect_f64", box_a, box_b)


def calculate_distance(box_a, box_b):
    """The tracker's alternative association metric (use_iou=False, iouTracke_cal.py:135-138)   [calc_performance.py:34-51]"""
    return _pairwise("fdt_calculate_distance_f64", box_a, box_b)


def calculate_iou(box_a, box_b):
    """Pairwise jaccard overlap of [A,4] and [B,4] corner boxes -> ndarray [A,B]   [calc_performance.py:54-74]"""
    return _pairwise("fdt_calculate_iou_f64", box_a, box_b)


def calc_pr(predict, truth, iou_thresh=0.5):
    """predict [P,5] rows [x1,y1,x2,y2,score], truth [T,4] rows [x,y,w,h]
    -> (ndarray [2,P] = [[IoU-matched flags], [scores]], truth_num)   [calc_performance.py:77-92]"""
    dev = _lib.require_cuda()
    predict = np.ascontiguousarray(predict, dtype=np.float64)
    truth = np.ascontiguousarray(truth, dtype=np.float64).reshape(-1, 4)
    p = torch.as_tensor(predict).to(dev)
    t = torch.as_tensor(truth).to(dev)
    tf = torch.zeros(predict.shape[0], dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().fdt_calc_pr(_lib.ptr(p), predict.shape[0], predict.shape[1], _lib.ptr(t), truth.shape[0],
                                          float(iou_thresh), _lib.ptr(tf), _lib.stream_ptr()))
    return np.vstack((tf.cpu().numpy(), predict[:, 4])), truth.shape[0]

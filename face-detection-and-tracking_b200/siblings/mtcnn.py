"""MTCNN's NMS helpers on the GPU: MTCNN/mtcnn/core/utils.py:62-113 `nms` and MTCNN/mtcnn/core/nms.py:4-40 `torch_nms`.

Same signatures and return type (a python list of indices into `dets`, descending score).  The reference runs numpy in the
dtype of `dets` (float64 in its pipeline): float64 `dets` are processed in float64 (fdt_nms_variant_f64), float32 `dets` in fp32.
Sort ties are unspecified in the reference (argsort is unstable): the higher index comes first here."""
from __future__ import annotations

import numpy as np

from .. import _lib
from ._nms import nms_variant


def _mode_flag(mode):
    if mode == "Union":
        return _lib.NMS_SUMFIRST
    if mode == "Minimum":
        return _lib.NMS_MINIMUM
    raise UnboundLocalError("local variable 'ovr' referenced before assignment")     # what the reference does for another mode


def nms(dets, thresh, mode="Union"):
    """dets [[x1, y1, x2, y2, score]] -> indexes to keep (utils.py:62-113): survive iff overlap < thresh."""
    dets = np.asarray(dets)
    return nms_variant(dets[:, :4], dets[:, 4], thresh, _mode_flag(mode), keep_dtype=True).cpu().tolist()


def torch_nms(dets, thresh, mode="Union"):
    """core/nms.py:4-40: "+ 1" pixel areas, survive iff overlap <= thresh."""
    dets = np.asarray(dets)
    return nms_variant(dets[:, :4], dets[:, 4], thresh, _mode_flag(mode) | _lib.NMS_PLUS1 | _lib.NMS_LE, keep_dtype=True).cpu().tolist()

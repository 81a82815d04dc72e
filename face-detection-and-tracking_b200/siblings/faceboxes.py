"""FaceBoxes' DataEncoder (FACEBOX/encoderl.py) on the GPU kernels: same class, method names and return types.

  decode_np  (:308-325)  threshold + decode + nms_np in three launches, no host round trip before the final read-out
  nms_np     (:218-266)  numpy NMS, "Union" / "Minimum", survive iff overlap < threshold
  nms        (:268-306)  torch NMS, survive iff overlap <= threshold
  encode     (:158-215)  matching + encoding = layers.box_utils.match_ensure_max_prior with conf[best prior of each face] = 1
The 21,824 default boxes (:21-47) are a constructor-time table built with the reference's python-float arithmetic."""
from __future__ import annotations

import itertools

import numpy as np
import torch

from .. import _lib
from ._nms import nms_variant


class DataEncoder:
    def __init__(self):
        scale = 1024.
        steps = [s / scale for s in (32, 64, 128)]
        sizes = [s / scale for s in (32, 256, 512)]
        aspect_ratios = ((1, 2, 4), (1,), (1,))
        feature_map_sizes = (32, 16, 8)
        density = [[-3, -1, 1, 3], [-1, 1], [0]]
        boxes = []
        for i, fmsize in enumerate(feature_map_sizes):
            for h, w in itertools.product(range(fmsize), repeat=2):
                cx = (w + 0.5) * steps[i]
                cy = (h + 0.5) * steps[i]
                s = sizes[i]
                for j, ar in enumerate(aspect_ratios[i]):
                    if i == 0:
                        for dx, dy in itertools.product(density[j], repeat=2):
                            boxes.append((cx + dx / 8. * s * ar, cy + dy / 8. * s * ar, s * ar, s * ar))
                    else:
                        boxes.append((cx, cy, s * ar, s * ar))
        self.default_boxes = torch.Tensor(boxes)
        self.default_boxes_np = self.default_boxes.numpy()
        self._dev = {}

    def _default_boxes_on(self, dev):
        t = self._dev.get(dev.index)
        if t is None:
            t = self._dev[dev.index] = self.default_boxes.to(dev).contiguous()
        return t

    # ------------------------------------------------------------------ NMS
    @staticmethod
    def nms_np(bboxes, scores, threshold=0.5, mode="Union"):
        flag = {"Union": _lib.NMS_SUMFIRST, "Minimum": _lib.NMS_MINIMUM}[mode]
        return nms_variant(bboxes, scores, threshold, flag).cpu().tolist()

    @staticmethod
    def nms(bboxes, scores, threshold=0.5):
        return nms_variant(bboxes, scores, threshold, _lib.NMS_SUMFIRST | _lib.NMS_LE).cpu()

    # ------------------------------------------------------------------ decode
    def decode_np(self, loc, conf, conf_thres=0.35, return_index=False):
        """loc [21824,4], conf [21824,2] -> (boxes[keep], scores[keep]) numpy arrays, NMS "Union" at 0.5."""
        dev = _lib.require_cuda()
        loc_d = _lib.dev_f32(torch.as_tensor(loc), dev).view(-1, 4)
        conf_d = _lib.dev_f32(torch.as_tensor(conf), dev).view(-1, 2)
        N = loc_d.shape[0]
        db = self._default_boxes_on(dev)
        if db.shape[0] != N:
            raise ValueError(f"decode_np: {N} predictions for {db.shape[0]} default boxes")
        boxes = torch.empty((N, 4), dtype=torch.float32, device=dev)
        keep = torch.zeros(N, dtype=torch.int64, device=dev)
        count = torch.zeros(1, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            L = _lib.lib()
            st = _lib.stream_ptr()
            _lib.check(L.fdt_facebox_decode(_lib.ptr(loc_d), _lib.ptr(db), N, 0.1, 0.2, _lib.ptr(boxes), st))
            ws = _lib.workspace(L.fdt_threshold_nms_workspace_bytes(N), dev, "fbx")
            _lib.check(L.fdt_threshold_nms(_lib.ptr(boxes), _lib.ptr(conf_d), N, float(conf_thres), 0.5, _lib.NMS_SUMFIRST,
                                           _lib.ptr(keep), _lib.ptr(count), _lib.ptr(ws), ws.numel(), st))
        c = int(count.item())
        if c < 0:
            raise NotImplementedError(f"decode_np: more than {_lib.MAX_NMS_TOP_K} boxes above conf_thres")
        k = keep[:c]
        out = (boxes[k].cpu().numpy(), conf_d[:, 1][k].cpu().numpy())
        return out + (k.cpu().numpy(),) if return_index else out

    # ------------------------------------------------------------------ encode
    def encode(self, boxes, classes, threshold=0.35):
        """boxes [num_obj,4] corner form, classes [num_obj] -> (loc [21824,4], conf [21824]) (encoderl.py:158-215)."""
        dev = _lib.require_cuda()
        src = boxes.device
        b = _lib.dev_f32(boxes, dev).view(-1, 4)
        if b.shape[0] == 0:
            raise IndexError("max(): Expected reduction dim to have non-zero size.")
        cls = classes.to(dev).view(-1)
        db = self._default_boxes_on(dev)
        N, G = db.shape[0], b.shape[0]
        gt = torch.cat([b, (cls - 1).to(torch.float32).unsqueeze(1)], 1).contiguous()       # match_* store label + 1
        off = torch.tensor([0, G], dtype=torch.int64, device=dev)
        loc_t = torch.empty((1, N, 4), dtype=torch.float32, device=dev)
        conf_t = torch.empty((1, N), dtype=torch.int64, device=dev)
        bto = torch.empty((1, N), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            L = _lib.lib()
            ws = _lib.workspace(L.fdt_match_workspace_bytes(1, N, G), dev, "match")
            _lib.check(L.fdt_match_encode(_lib.ptr(db), _lib.ptr(gt), _lib.ptr(off), G, 1, N, float(threshold), 0.1, 0.2, 1,
                                          _lib.ptr(loc_t), _lib.ptr(conf_t), None, _lib.ptr(bto), _lib.ptr(ws), ws.numel(),
                                          _lib.stream_ptr()))
        bto = bto[0]
        loc, conf = loc_t[0], conf_t[0]
        conf[bto == 2.0] = 1                                                   # :201 conf[max_iou_index] = 1
        if bool((loc[:, 2:].abs() > 10000).any()):                             # :195-200 (the reference raises NameError inf_error)
            raise RuntimeError("encode: inf_flag has true (a zero-width or zero-height box)")
        return loc.to(src), conf.to(classes.dtype).to(src)

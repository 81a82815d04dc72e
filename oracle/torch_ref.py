"""Torch restatement of the reference's Detect / nms python loops -- TEST / BENCH INFRASTRUCTURE, not product code.

Purpose: BASELINE.md section 5.4 asks for the reference's algorithm timed on torch-CUDA tensors on the same B200 (its
intended deployment: a python loop of small ATen kernels with a host sync per iteration).  /root/reference does not exist on
the GPU box, so this module restates layers/box_utils.py:238-258 (decode), :275-340 (nms) and
layers/functions/detection.py:34-84 (Detect.__call__) op by op in torch; tests check it against the C oracle.
Only tests/ and bench_extra.py import it."""
from __future__ import annotations

import torch


def decode(loc, priors, variances):
    """box_utils.py:238-258"""
    cxcy = priors[:, :2] + loc[:, :2] * variances[0] * priors[:, 2:]
    wh = priors[:, 2:] * torch.exp(loc[:, 2:] * variances[1])
    x1y1 = cxcy - wh / 2
    return torch.cat((x1y1, wh + x1y1), 1)


def nms(boxes, scores, overlap=0.5, top_k=200):
    """box_utils.py:275-340: sort ascending, keep the last top_k, pop the best, drop what overlaps it, repeat."""
    keep = scores.new_zeros(scores.size(0)).long()
    if boxes.numel() == 0:
        return keep, 0
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    area = (x2 - x1) * (y2 - y1)
    _, idx = scores.sort(0)
    idx = idx[-top_k:]
    count = 0
    while idx.numel() > 0:
        i = idx[-1]
        keep[count] = i
        count += 1
        if idx.size(0) == 1:
            break
        idx = idx[:-1]
        xx1 = x1[idx].clamp(min=float(x1[i]))        # float(): the host sync the reference pays through index_select + clamp
        yy1 = y1[idx].clamp(min=float(y1[i]))
        xx2 = x2[idx].clamp(max=float(x2[i]))
        yy2 = y2[idx].clamp(max=float(y2[i]))
        w = (xx2 - xx1).clamp(min=0.0)
        h = (yy2 - yy1).clamp(min=0.0)
        inter = w * h
        union = (area[idx] - inter) + area[i]
        idx = idx[(inter / union).lt(overlap)]
    return keep, count


class Detect:
    """detection.py:9-84"""

    def __init__(self, num_classes, bkg_label, top_k, conf_thresh, nms_thresh, variance=(0.1, 0.2)):
        self.num_classes, self.background_label, self.top_k = num_classes, bkg_label, top_k
        self.conf_thresh, self.nms_thresh, self.variance, self.nms_top_k = conf_thresh, nms_thresh, variance, 5000

    def __call__(self, loc_data, conf_data, prior_data):
        num = loc_data.size(0)
        output = torch.zeros(num, self.num_classes, self.top_k, 5, device=loc_data.device)
        conf_preds = conf_data.view(num, prior_data.size(0), self.num_classes).transpose(2, 1)
        for i in range(num):
            decoded = decode(loc_data[i], prior_data, self.variance)
            for cl in range(1, self.num_classes):
                c_mask = conf_preds[i][cl].gt(self.conf_thresh)
                scores = conf_preds[i][cl][c_mask]
                if scores.numel() <= 1:                 # zero candidates, or the reference's 0-d `continue` quirk (:66-72)
                    continue
                boxes = decoded[c_mask.unsqueeze(1).expand_as(decoded)].view(-1, 4)
                ids, count = nms(boxes, scores, self.nms_thresh, self.nms_top_k)
                count = min(count, self.top_k)
                output[i, cl, :count] = torch.cat((scores[ids[:count]].unsqueeze(1), boxes[ids[:count]]), 1)
        return output

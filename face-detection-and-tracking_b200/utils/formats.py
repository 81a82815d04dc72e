"""On-disk formats of the reference around the box pipeline.  Host-side file handling only (no kernels):

  PR data    My_test.py:166-171 writes, draw_curve/draw_pr_roc.py:5-34 reads: float64 array [2, M+1]; row 0 = matched flags,
             row 1 = scores, columns sorted by descending score, last column = [0, truth_num].
  annotation data/widerface.py:88-91 and utils/data_collector.py:23-26, :43-51: one image per line,
             "<path> <num> x y w h x y w h ...";  AnnotationTransform (data/widerface.py:35-63) -> normalised corner boxes.
  tracks     fdt_b200.tracker.save_tracks / load_tracks (iouTracke_cal.py:177, iouTracke_display.py:29).
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------------------------- PR data
def new_pr_accumulator():
    """tf_conf = np.array([[], []])   [My_test.py:105]"""
    return np.array([[], []])


def accumulate_pr(tf_conf, tf_conf_image):
    """tf_conf = np.hstack((tf_conf, tf_conf_))   [My_test.py:161]; tf_conf_image is calc_pr(...)[0]."""
    return np.hstack((tf_conf, tf_conf_image))


def save_pr_data(path, tf_conf, truth_num):
    """Sort the columns by descending score and append [0, truth_num]   [My_test.py:166-171]."""
    tf_conf = np.asarray(tf_conf)
    tf_conf = tf_conf[:, np.argsort(tf_conf[1, :])[::-1]]
    data = np.hstack((tf_conf, [[0], [truth_num]]))
    np.save(path, data)
    return data


def load_pr_data(path):
    """-> (tf_conf [2, M], truth_num)   [draw_pr_roc.py:27-30]"""
    data = np.load(path)
    return data[:, :-1], data[1, -1]


def gen_tp_fp(tf_conf):
    """true_pos[i] = number of matched detections among the i+1 best, false_pos[i] = i + 1 - true_pos[i]   [draw_pr_roc.py:5-19]"""
    _, M = tf_conf.shape
    true_pos = np.cumsum(tf_conf[0, :] != 0).astype(np.float64)
    false_pos = np.arange(1, M + 1, dtype=np.float64) - true_pos
    return true_pos, false_pos


def pr_roc(data):
    """Full [2, M+1] array -> ((recall, precision), (false_pos, recall))   [draw_pr_roc.py:27-34]"""
    truth_num = data[1, -1]
    tp, fp = gen_tp_fp(data[:, :-1])
    recall = tp / truth_num
    precision = tp / (tp + fp)
    return (recall, precision), (fp, recall)


# ----------------------------------------------------------------------------------------------- annotation lines
def read_annotation_file(anno_file):
    """-> (ids, annotation): image paths and the remaining fields of each line, as strings   [data/widerface.py:88-91]"""
    ids, annotation = [], []
    with open(anno_file, 'r') as f:
        for line in f:
            fields = line.strip().split()
            if not fields:
                continue
            ids.append(fields[0])
            annotation.append(fields[1:])
    return ids, annotation


def annotation_to_pixel_boxes(target):
    """["num", x, y, w, h, ...] -> int32 [num, 4] rows [x, y, w, h] (calc_pr's `truth`)   [utils/data_collector.py:46-51]"""
    num = int(target[0])
    return np.array(target[1:1 + 4 * num]).astype(np.int32).reshape(num, 4)


class AnnotationTransform(object):
    """widerface annotation -> [[xmin, ymin, xmax, ymax, label_ind], ...] normalised by the image size   [data/widerface.py:19-63].
    Boxes with zero width or height are dropped; a negative width (else: a negative height) swaps that pair of corners."""

    def __call__(self, target, width, height):
        num = int(target[0])
        res = []
        for i in range(num):
            xmin = int(target[1 + i * 4])
            ymin = int(target[2 + i * 4])
            xmax = int(target[3 + i * 4]) + xmin
            ymax = int(target[4 + i * 4]) + ymin
            if int(target[3 + i * 4]) == 0 or int(target[4 + i * 4]) == 0:
                continue
            elif int(target[3 + i * 4]) < 0:
                xmin, xmax = xmax, xmin
            elif int(target[4 + i * 4]) < 0:
                ymin, ymax = ymax, ymin
            res.append([xmin / float(width), ymin / float(height), xmax / float(width), ymax / float(height), 0])
        return res


def annotation_to_targets(annotation, sizes, device=None):
    """Annotation field lists + (width, height) per image -> the `targets` list MultiBoxLoss.forward takes: one float32
    tensor [G_i, 5] per image (zero-box images give [0, 5])."""
    import torch
    tr = AnnotationTransform()
    out = []
    for target, (w, h) in zip(annotation, sizes):
        rows = tr(target, w, h)
        t = torch.tensor(rows, dtype=torch.float32).reshape(-1, 5)
        out.append(t.to(device) if device is not None else t)
    return out

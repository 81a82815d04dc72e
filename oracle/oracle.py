"""ctypes front-end of the CPU ORACLE (oracle/fdt_oracle.c).

TEST INFRASTRUCTURE -- only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  The product package never does.

Function names and argument meaning follow the reference (layers/box_utils.py,
layers/functions/*.py, layers/modules/multibox_loss.py, utils/calc_performance.py,
iouTracke_cal.py); arrays are numpy, fp32 unless stated.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from math import sqrt

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfdt_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with the PATH gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "fdt_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", _HERE, "-B", "libfdt_oracle.so"], check=True, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_nms.restype = C.c_int64
        _lib.orc_iou_track.restype = C.c_int64
        _lib.orc_match.restype = C.c_int
    return _lib


def max_threads() -> int:
    return int(lib().orc_max_threads())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ---------------------------------------------------------------- PriorBoxLayer (prior_box.py:9-44)
class PriorBoxLayer:
    def __init__(self, width, height, stride=(4, 8, 16, 32, 64, 128), box=(16, 32, 64, 128, 256, 512),
                 scale=(1, 1, 1, 1, 1, 1), aspect_ratios=([], [], [], [], [], [])):
        self.width, self.height = width, height
        self.stride, self.box, self.scales, self.aspect_ratios = stride, box, scale, aspect_ratios

    def __call__(self, prior_idx, f_width, f_height):
        ns = int(self.scales[prior_idx])
        ars = list(self.aspect_ratios[prior_idx])
        bs = np.array([(2 ** (1 / 3)) ** s for s in range(ns)], dtype=np.float64)      # prior_box.py:33
        sa = np.array([sqrt(ar) for ar in ars], dtype=np.float64)                       # prior_box.py:41
        out = np.empty((f_height * f_width * ns * (1 + len(ars)), 4), dtype=np.float32)
        lib().orc_priorbox(C.c_double(self.width), C.c_double(self.height),
                           C.c_double(self.stride[prior_idx]), C.c_double(self.box[prior_idx]),
                           C.c_int(ns), _p(bs), C.c_int(len(ars)), _p(sa),
                           C.c_int(f_width), C.c_int(f_height), _p(out))
        return out


# ---------------------------------------------------------------- box_utils
def point_form(boxes):
    b = _f32(boxes); out = np.empty_like(b)
    lib().orc_point_form(_p(b), C.c_int64(b.shape[0]), _p(out)); return out


def center_size(boxes):
    b = _f32(boxes); out = np.empty_like(b)
    lib().orc_center_size(_p(b), C.c_int64(b.shape[0]), _p(out)); return out


def intersect(box_a, box_b):
    a, b = _f32(box_a), _f32(box_b); out = np.empty((a.shape[0], b.shape[0]), np.float32)
    lib().orc_intersect(_p(a), C.c_int64(a.shape[0]), _p(b), C.c_int64(b.shape[0]), _p(out)); return out


def calculate_iou(box_a, box_b):
    a, b = _f32(box_a), _f32(box_b); out = np.empty((a.shape[0], b.shape[0]), np.float32)
    lib().orc_calculate_iou(_p(a), C.c_int64(a.shape[0]), _p(b), C.c_int64(b.shape[0]), _p(out)); return out


def encode(matched, priors, variances):
    m, p = _f32(matched), _f32(priors); out = np.empty_like(m)
    lib().orc_encode(_p(m), _p(p), C.c_int64(m.shape[0]), C.c_float(variances[0]), C.c_float(variances[1]), _p(out))
    return out


def decode(loc, priors, variances):
    l, p = _f32(loc), _f32(priors); out = np.empty_like(l)
    lib().orc_decode(_p(l), _p(p), C.c_int64(l.shape[0]), C.c_float(variances[0]), C.c_float(variances[1]), _p(out))
    return out


def log_sum_exp(x):
    x = _f32(x); out = np.empty((x.shape[0], 1), np.float32)
    lib().orc_log_sum_exp(_p(x), C.c_int64(x.shape[0]), C.c_int(x.shape[1]), _p(out)); return out


def nms(boxes, scores, overlap=0.5, top_k=200):
    """-> (keep int64[n] zero-padded, count)   (box_utils.py:275-340)"""
    b, s = _f32(boxes).reshape(-1, 4), _f32(scores).reshape(-1)
    keep = np.zeros(s.shape[0], np.int64)
    cnt = lib().orc_nms(_p(b), _p(s), C.c_int64(s.shape[0]), C.c_float(overlap), C.c_int64(top_k), _p(keep))
    return keep, int(cnt)


def match(bipartite, threshold, truth_loc, priors, variances, truth_conf):
    """match_ensure_max_prior (bipartite=True) / match_default for one image.
    -> loc_t[N,4], conf_t[N] int64, best_truth_idx[N] int64, best_truth_overlap[N]"""
    t, p, lab = _f32(truth_loc).reshape(-1, 4), _f32(priors), _f32(truth_conf).reshape(-1)
    N = p.shape[0]
    loc_t = np.empty((N, 4), np.float32); conf_t = np.empty(N, np.int64)
    bti = np.empty(N, np.int64); bto = np.empty(N, np.float32)
    rc = lib().orc_match(C.c_int(int(bool(bipartite))), C.c_float(threshold), _p(t), _p(lab), C.c_int64(t.shape[0]),
                         _p(p), C.c_int64(N), C.c_float(variances[0]), C.c_float(variances[1]),
                         _p(loc_t), _p(conf_t), _p(bti), _p(bto))
    if rc != 0:
        raise IndexError("max(): Expected reduction dim 0 to have non-zero size (G == 0; reference raises too)")
    return loc_t, conf_t, bti, bto


def hard_negative_mine(loss_c, pos, negpos_ratio, n_threads=0):
    lc = _f32(loss_c); ps = np.ascontiguousarray(pos, dtype=np.uint8)
    B, N = lc.shape
    neg = np.empty((B, N), np.uint8)
    lib().orc_hard_negative_mine(_p(lc), _p(ps), C.c_int(B), C.c_int64(N), C.c_int(negpos_ratio), _p(neg), C.c_int(n_threads))
    return neg.astype(bool)


def pack_targets(targets):
    """list[B] of [G_i,5] -> (gt[sum G,5] fp32, off[B+1] int64)"""
    off = np.zeros(len(targets) + 1, np.int64)
    for i, t in enumerate(targets):
        off[i + 1] = off[i] + (0 if t is None else int(np.asarray(t).reshape(-1, 5).shape[0]))
    gt = np.zeros((max(int(off[-1]), 1), 5), np.float32)
    for i, t in enumerate(targets):
        if off[i + 1] > off[i]:
            gt[off[i]:off[i + 1]] = np.asarray(t, dtype=np.float32).reshape(-1, 5)
    return gt, off


def multibox_loss(loc, conf, priors, targets, threshold=0.35, negpos_ratio=3, bipartite=False,
                  variances=(0.1, 0.2), n_threads=0, want_aux=True):
    """MultiBoxLoss.forward (multibox_loss.py:48-136).  -> dict(loss_l, loss_c, loc_t, conf_t, loss_c_all, neg)"""
    l, c, p = _f32(loc), _f32(conf), _f32(priors)
    B, N, Cn = c.shape
    gt, off = pack_targets(targets)
    losses = np.zeros(2, np.float32)
    loc_t = np.empty((B, N, 4), np.float32) if want_aux else None
    conf_t = np.empty((B, N), np.int64) if want_aux else None
    lca = np.empty((B, N), np.float32) if want_aux else None
    neg = np.empty((B, N), np.uint8) if want_aux else None
    lib().orc_multibox_loss(_p(l), _p(c), _p(p), _p(gt), _p(off), C.c_int(B), C.c_int64(N), C.c_int(Cn),
                            C.c_float(threshold), C.c_int(negpos_ratio), C.c_int(int(bool(bipartite))),
                            C.c_float(variances[0]), C.c_float(variances[1]),
                            _p(losses), _p(loc_t), _p(conf_t), _p(lca), _p(neg), C.c_int(n_threads))
    return dict(loss_l=float(losses[0]), loss_c=float(losses[1]), loc_t=loc_t, conf_t=conf_t, loss_c_all=lca,
                neg=None if neg is None else neg.astype(bool))


# ---------------------------------------------------------------- Detect (detection.py:9-84)
class Detect:
    def __init__(self, num_classes, bkg_label, top_k, conf_thresh, nms_thresh):
        self.num_classes, self.background_label, self.top_k = num_classes, bkg_label, top_k
        self.nms_thresh = nms_thresh
        if nms_thresh <= 0:
            raise ValueError('nms_threshold must be non negative.')
        self.conf_thresh = conf_thresh
        self.variance = [0.1, 0.2]          # data/config.py:17
        self.nms_top_k = 5000               # detection.py:32
        self.n_threads = 0
        self.early_exit = False

    def __call__(self, loc_data, conf_data, prior_data, return_aux=False):
        p = _f32(prior_data); N = p.shape[0]
        l = _f32(loc_data).reshape(-1, N, 4); B = l.shape[0]
        c = _f32(conf_data).reshape(B, N, self.num_classes)
        out = np.empty((B, self.num_classes, self.top_k, 5), np.float32)
        counts = np.empty((B, self.num_classes), np.int32)
        kept = np.empty((B, self.num_classes, self.top_k), np.int64)
        lib().orc_detect(_p(l), _p(c), _p(p), C.c_int(B), C.c_int64(N), C.c_int(self.num_classes),
                         C.c_int(self.top_k), C.c_int(self.nms_top_k), C.c_float(self.conf_thresh),
                         C.c_float(self.nms_thresh), C.c_float(self.variance[0]), C.c_float(self.variance[1]),
                         _p(out), _p(counts), _p(kept), C.c_int(int(self.early_exit)), C.c_int(self.n_threads))
        return (out, counts, kept) if return_aux else out


# ---------------------------------------------------------------- utils.calc_performance (float64)
def calculate_iou_f64(box_a, box_b):
    a = np.ascontiguousarray(box_a, dtype=np.float64); b = np.ascontiguousarray(box_b, dtype=np.float64)
    out = np.empty((a.shape[0], b.shape[0]), np.float64)
    lib().orc_calculate_iou_f64(_p(a), C.c_int64(a.shape[0]), _p(b), C.c_int64(b.shape[0]), _p(out)); return out


def intersect_f64(box_a, box_b):
    a = np.ascontiguousarray(box_a, dtype=np.float64); b = np.ascontiguousarray(box_b, dtype=np.float64)
    out = np.empty((a.shape[0], b.shape[0]), np.float64)
    lib().orc_intersect_f64(_p(a), C.c_int64(a.shape[0]), _p(b), C.c_int64(b.shape[0]), _p(out)); return out


def calculate_distance_f64(box_a, box_b):
    a = np.ascontiguousarray(box_a, dtype=np.float64); b = np.ascontiguousarray(box_b, dtype=np.float64)
    out = np.empty((a.shape[0], b.shape[0]), np.float64)
    lib().orc_calculate_distance_f64(_p(a), C.c_int64(a.shape[0]), _p(b), C.c_int64(b.shape[0]), _p(out)); return out


def calc_pr(predict, truth, iou_thresh=0.5):
    """-> (ndarray [2, P] = [tf; score], truth_num)   (calc_performance.py:77-92)"""
    p = np.ascontiguousarray(predict, dtype=np.float64); t = np.ascontiguousarray(truth, dtype=np.float64)
    tf = np.zeros(p.shape[0], np.int32)
    lib().orc_calc_pr.restype = C.c_int64
    n = lib().orc_calc_pr(_p(p), C.c_int64(p.shape[0]), C.c_int(p.shape[1]), _p(t), C.c_int64(t.shape[0]), C.c_double(iou_thresh), _p(tf))
    return np.vstack((tf, p[:, 4])), int(n)


# ---------------------------------------------------------------- tracker (iouTracke_cal.py:126-155,174-176)
def iou_track_raw(frames, sigma_iou=0.4, sigma_h=0.6, t_min=5, use_iou=True, sigma_dis=8):
    """frames: list of [D_f,5] arrays.  -> (dets[total,5] f64, track_off, track_dets, track_start, track_max)
    use_iou=False: association by calculate_distance / argmin / < sigma_dis (iouTracke_cal.py:135-138)."""
    off = np.zeros(len(frames) + 1, np.int64)
    for i, f in enumerate(frames):
        off[i + 1] = off[i] + np.asarray(f).reshape(-1, 5).shape[0]
    total = int(off[-1])
    dets = np.zeros((max(total, 1), 5), np.float64)
    for i, f in enumerate(frames):
        if off[i + 1] > off[i]:
            dets[off[i]:off[i + 1]] = np.asarray(f, dtype=np.float64).reshape(-1, 5)
    t_off = np.zeros(total + 2, np.int64); t_dets = np.zeros(max(total, 1), np.int64)
    t_start = np.zeros(total + 1, np.int64); t_max = np.zeros(total + 1, np.float64)
    lib().orc_iou_track_metric.restype = C.c_int64
    T = lib().orc_iou_track_metric(_p(dets), _p(off), C.c_int64(len(frames)), C.c_int(0 if use_iou else 1),
                                   C.c_double(sigma_iou if use_iou else sigma_dis), C.c_double(sigma_h),
                                   C.c_int64(t_min), _p(t_off), _p(t_dets), _p(t_start), _p(t_max))
    return dets, t_off[:T + 1], t_dets[:int(t_off[T])], t_start[:T], t_max[:T]


def iou_track(frames, sigma_iou=0.4, sigma_h=0.6, t_min=5, use_iou=True, sigma_dis=8):
    """-> tracks_finished: list of {'bboxes': [[x1,y1,x2,y2],...], 'max_score': float, 'start_frame': int}"""
    dets, t_off, t_dets, t_start, t_max = iou_track_raw(frames, sigma_iou, sigma_h, t_min, use_iou, sigma_dis)
    out = []
    for t in range(len(t_start)):
        rows = t_dets[t_off[t]:t_off[t + 1]]
        out.append({'bboxes': dets[rows, :4].tolist(), 'max_score': float(t_max[t]), 'start_frame': int(t_start[t])})
    return out


def detections_to_frames(detections, width, height, thresh=0.4, shrink=1.0):
    """iouTracke_cal.py:55-84 (detect_face after the network call) for a clip: detections [F, C, top_k, 5] float32 -> list of F arrays
    [n_f, 5]: class by class the leading rows with score >= thresh (:61-68), boxes * (w, h, w, h) in float32 (:64), / shrink in float32
    (:76-79: numpy keeps float32 against a python scalar), score appended (:80); the float64 dummy [[0,0,0,0,0.4]] if none (:73-74)."""
    det = np.asarray(detections, dtype=np.float32)
    scale = np.array([width, height, width, height], np.float32)
    out = []
    for f in range(det.shape[0]):
        rows = []
        for i in range(det.shape[1]):
            j = 0
            while j < det.shape[2] and det[f, i, j, 0] >= np.float32(thresh):
                pt = det[f, i, j, 1:] * scale
                rows.append(np.concatenate([pt / np.float32(shrink), det[f, i, j, :1]]))
                j += 1
        out.append(np.array(rows, np.float32).reshape(-1, 5) if rows else np.array([[0, 0, 0, 0, 0.4]]))
    return out


# ---------------------------------------------------------------- head post-processing (pyramid.py:291-309, 331-332)
def heads_to_loc_conf(loc_maps, conf_maps, neg_max=None, softmax=True):
    """loc_maps / conf_maps: per-level [B,4,H,W] fp32 arrays (NCHW) -> (loc[B,N,4], conf[B,N,2]).
    neg_max[l] truthy: neg = max(ch0..2), pos = ch3 (level 0 in every model); default = (1, 0, 0, ...)."""
    L = len(conf_maps)
    conf_maps = [_f32(m) for m in conf_maps]
    loc_maps = [_f32(m) for m in loc_maps]
    B = conf_maps[0].shape[0]
    fh = (C.c_int * L)(*[m.shape[2] for m in conf_maps])
    fw = (C.c_int * L)(*[m.shape[3] for m in conf_maps])
    nm = (C.c_int * L)(*([1] + [0] * (L - 1) if neg_max is None else [int(bool(v)) for v in neg_max]))
    N = sum(m.shape[2] * m.shape[3] for m in conf_maps)
    loc = np.empty((B, N, 4), dtype=np.float32)
    conf = np.empty((B, N, 2), dtype=np.float32)
    lp = (C.c_void_p * L)(*[m.ctypes.data for m in loc_maps])
    cp = (C.c_void_p * L)(*[m.ctypes.data for m in conf_maps])
    lib().orc_heads_to_loc_conf(lp, cp, fh, fw, nm, C.c_int(L), C.c_int(B), C.c_int(1 if softmax else 0), _p(loc), _p(conf))
    return loc, conf


# ---------------------------------------------------------------- sibling NMS / decode implementations
NMS_SUMFIRST, NMS_MINIMUM, NMS_PLUS1, NMS_LE = 1, 2, 4, 8


def nms_variant(boxes, scores, thresh, flags):
    """-> kept indices (int64[count]) in keep order; see orc_nms_variant for the flag meaning."""
    b = _f32(boxes).reshape(-1, 4); s = _f32(scores).reshape(-1)
    keep = np.zeros(s.shape[0], dtype=np.int64)
    lib().orc_nms_variant.restype = C.c_int64
    c = lib().orc_nms_variant(_p(b), _p(s), C.c_int64(s.shape[0]), C.c_float(thresh), C.c_int(flags), _p(keep))
    return keep[:int(c)]


def nms_variant_f64(boxes, scores, thresh, flags):
    """float64 form (MTCNN's nms on float64 dets) -> kept indices in keep order."""
    b = np.ascontiguousarray(boxes, dtype=np.float64).reshape(-1, 4); s = np.ascontiguousarray(scores, dtype=np.float64).reshape(-1)
    keep = np.zeros(s.shape[0], dtype=np.int64)
    lib().orc_nms_variant_f64.restype = C.c_int64
    c = lib().orc_nms_variant_f64(_p(b), _p(s), C.c_int64(s.shape[0]), C.c_double(thresh), C.c_int(flags), _p(keep))
    return keep[:int(c)]


def facebox_default_boxes():
    """FACEBOX/encoderl.py:21-46: python-float arithmetic, one rounding to fp32 at torch.Tensor(boxes)."""
    import itertools
    scale = 1024.
    steps = [s / scale for s in (32, 64, 128)]
    sizes = [s / scale for s in (32, 256, 512)]
    aspect_ratios = ((1, 2, 4), (1,), (1,))
    feature_map_sizes = (32, 16, 8)
    density = [[-3, -1, 1, 3], [-1, 1], [0]]
    boxes = []
    for i in range(len(feature_map_sizes)):
        fmsize = feature_map_sizes[i]
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
    return np.array(boxes, dtype=np.float64).astype(np.float32)


def facebox_decode(loc, default_boxes):
    l = _f32(loc).reshape(-1, 4); d = _f32(default_boxes).reshape(-1, 4)
    out = np.empty_like(l)
    lib().orc_facebox_decode(_p(l), _p(d), C.c_int64(l.shape[0]), C.c_float(0.1), C.c_float(0.2), _p(out))
    return out


def facebox_decode_np(loc, conf, default_boxes, conf_thres=0.35, nms_thresh=0.5):
    """DataEncoder.decode_np (encoderl.py:308-325) -> (boxes[keep], scores[keep])."""
    score = _f32(conf)[:, 1]
    ids = np.where(score > np.float32(conf_thres))[0]
    boxes = facebox_decode(_f32(loc)[ids], _f32(default_boxes)[ids])
    keep = nms_variant(boxes, score[ids], nms_thresh, NMS_SUMFIRST)
    return boxes[keep], score[ids][keep]

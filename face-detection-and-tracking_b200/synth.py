"""Seeded synthetic workloads for the box pipeline.

numpy-only and deterministic (PCG64 streams are stable across numpy versions), so tests, bench.py
and oracle/make_golden.py all see the same bits.  The reference ships no weights or videos; every
BASELINE.json config is "synthetic head outputs".
"""
from __future__ import annotations

import hashlib
from math import ceil

import numpy as np

STRIDES6 = (4, 8, 16, 32, 64, 128)
BOXES6 = (16, 32, 64, 128, 256, 512)


def feature_maps(width, height, strides=STRIDES6):
    """Feature-map sizes (f_w, f_h) per pyramid level; 640 -> 160,80,40,20,10,5 (data/config.py:5)."""
    return [(int(ceil(width / s)), int(ceil(height / s))) for s in strides]


def priors_numpy(width, height, strides=STRIDES6, boxes=BOXES6):
    """fp64 -> fp32 restatement of the prior set the models build (pyramid.py:270-285): centre-form
    [cx,cy,w,h], levels concatenated, y outer / x inner.  Generator-side helper only."""
    out = []
    for s, b, (fw, fh) in zip(strides, boxes, feature_maps(width, height, strides)):
        j = np.arange(fw, dtype=np.float64); i = np.arange(fh, dtype=np.float64)
        cx = np.broadcast_to(((j + 0.5) * s / width)[None, :], (fh, fw))
        cy = np.broadcast_to(((i + 0.5) * s / height)[:, None], (fh, fw))
        sx = np.full((fh, fw), b * 1.0 / width); sy = np.full((fh, fw), b * 1.0 / height)
        out.append(np.stack([cx, cy, sx, sy], -1).reshape(-1, 4))
    return np.concatenate(out, 0).astype(np.float32)


def _uniquify_candidates(score, thresh):
    """Nudge duplicated candidate scores by ulps so the sort order is unique."""
    for b in range(score.shape[0]):
        s = score[b]
        cand = np.nonzero(s > thresh)[0]
        seen = set()
        for p in cand[np.argsort(s[cand], kind="stable")]:
            v = s[p]
            while float(v) in seen:
                v = np.nextafter(v, np.float32(2.0), dtype=np.float32)
            seen.add(float(v)); s[p] = v
    return score


def detect_inputs(B, priors, seed, conf_thresh=0.05, mode="random"):
    """Synthetic PyramidBox head outputs: loc[B,N,4], conf[B,N,2] (post-softmax), fp32.

    mode="random":    loc ~ 0.5*N(0,1); logit gap d ~ N(-4.5, 2) -> ~22 % of priors above 0.05
                      (about 7.4 k candidates/image at N=34,125: exercises the 5000 truncation).
    mode="clustered": 1-50 synthetic faces/image; only priors with IoU > 0.35 to a face score high.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    N = priors.shape[0]
    loc = (rng.standard_normal((B, N, 4), dtype=np.float32) * np.float32(0.5)).astype(np.float32)
    if mode == "random":
        d = rng.standard_normal((B, N), dtype=np.float32) * np.float32(2.0) - np.float32(4.5)
    elif mode == "clustered":
        d = rng.standard_normal((B, N), dtype=np.float32) * np.float32(0.7) - np.float32(7.0)
        pf = np.concatenate([priors[:, :2] - priors[:, 2:] / 2, priors[:, :2] + priors[:, 2:] / 2], 1)
        for b in range(B):
            g = gt_boxes(int(rng.integers(1, 51)), rng)[:, :4]
            iou = _iou_np(g, pf).max(0)
            hit = iou > 0.35
            d[b, hit] = (rng.standard_normal(int(hit.sum()), dtype=np.float32) * np.float32(1.5) + np.float32(2.0))
            loc[b, hit] *= np.float32(0.3)
    else:
        raise ValueError(mode)
    s1 = (1.0 / (1.0 + np.exp(-d.astype(np.float64)))).astype(np.float32)
    s1 = _uniquify_candidates(s1, np.float32(conf_thresh))
    conf = np.stack([(np.float32(1.0) - s1), s1], -1).astype(np.float32)
    return loc, conf


def _iou_np(a, b):
    lt = np.maximum(a[:, None, :2], b[None, :, :2]); rb = np.minimum(a[:, None, 2:], b[None, :, 2:])
    wh = np.clip(rb - lt, 0, None); inter = wh[..., 0] * wh[..., 1]
    aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]); ab = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / (aa[:, None] + ab[None, :] - inter)


def gt_boxes(G, rng):
    """[G,5] fp32 rows [x1,y1,x2,y2,label=0]: centre U(0.1,0.9)^2, side U(0.01,0.2), w,h > 0."""
    c = rng.uniform(0.1, 0.9, (G, 2)); s = rng.uniform(0.01, 0.2, (G, 2))
    g = np.concatenate([c - s / 2, c + s / 2, np.zeros((G, 1))], 1).astype(np.float32)
    return g


def multibox_inputs(B, priors, seed, g_lo=0, g_hi=200):
    """loc ~ N(0,0.5^2), conf ~ N(0,1) raw logits [B,N,2], targets list[B] of [G_i,5], G_i ~ U{g_lo..g_hi}."""
    rng = np.random.Generator(np.random.PCG64(seed))
    N = priors.shape[0]
    loc = (rng.standard_normal((B, N, 4), dtype=np.float32) * np.float32(0.5)).astype(np.float32)
    conf = rng.standard_normal((B, N, 2), dtype=np.float32)
    targets = [gt_boxes(int(rng.integers(g_lo, g_hi + 1)), rng) for _ in range(B)]
    return loc, conf, targets


def tracker_frames(F, seed, d_lo=1, d_hi=300, n_objects=300, width=640.0, height=480.0, empty_every=1000,
                   sigma=1.5):
    """Per-frame float32 [D_f,5] rows [x1,y1,x2,y2,score] in pixels (what iouTracke_cal.detect_face
    returns, :70-84): n_objects persistent boxes doing pixel random walks, a random subset of D_f of
    them visible per frame; every `empty_every`-th frame is empty -> the reference's float64 dummy
    detection [[0,0,0,0,0.4]] (:73-74)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    side = rng.uniform(10, 60, (n_objects, 2))
    ctr = np.stack([rng.uniform(30, width - 30, n_objects), rng.uniform(30, height - 30, n_objects)], 1)
    frames = []
    for f in range(F):
        ctr = ctr + rng.normal(0, sigma, ctr.shape)
        ctr[:, 0] = np.clip(ctr[:, 0], 0, width); ctr[:, 1] = np.clip(ctr[:, 1], 0, height)
        if empty_every and (f + 1) % empty_every == 0:
            frames.append(np.array([[0, 0, 0, 0, 0.4]]))
            continue
        D = int(rng.integers(d_lo, min(d_hi, n_objects) + 1))
        vis = rng.permutation(n_objects)[:D]
        sc = rng.uniform(0.4, 1.0, D)
        det = np.concatenate([ctr[vis] - side[vis] / 2, ctr[vis] + side[vis] / 2, sc[:, None]], 1).astype(np.float32)
        frames.append(det)
    return frames


def clip_detections(F, top_k, seed):
    """Detect-shaped output of a clip, [F, 2, top_k, 5]: per frame a few faces that drift from frame to frame, rows in
    descending score order, then rows below the tracker's read-out threshold (0.4), then zero padding; some frames have no face."""
    rng = np.random.Generator(np.random.PCG64(seed))
    det = np.zeros((F, 2, top_k, 5), np.float32)
    n_obj = 9
    ctr = rng.uniform(0.15, 0.85, (n_obj, 2)); size = rng.uniform(0.04, 0.2, (n_obj, 2))
    for f in range(F):
        ctr += rng.normal(0, 0.004, ctr.shape)
        vis = rng.uniform(size=n_obj) < 0.8
        if f % 13 == 5:
            vis[:] = False
        rows = []
        for o in np.where(vis)[0]:
            c = ctr[o] + rng.normal(0, 0.002, 2); s = size[o] * rng.uniform(0.97, 1.03, 2)
            rows.append([rng.uniform(0.41, 0.999), c[0] - s[0] / 2, c[1] - s[1] / 2, c[0] + s[0] / 2, c[1] + s[1] / 2])
        for _ in range(int(rng.integers(0, 6))):                       # kept by NMS but below the tracker's 0.4
            c = rng.uniform(0.1, 0.9, 2); s = rng.uniform(0.03, 0.1, 2)
            rows.append([rng.uniform(0.05, 0.399), c[0] - s[0] / 2, c[1] - s[1] / 2, c[0] + s[0] / 2, c[1] + s[1] / 2])
        rows.sort(key=lambda r: -r[0])
        for j, r in enumerate(rows[:top_k]):
            det[f, 1, j] = np.array(r, np.float32)
    return det


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode()); h.update(str(a.shape).encode()); h.update(a.tobytes())
    return h.hexdigest()


def head_maps(B, width, height, seed, strides=STRIDES6):
    """Synthetic outputs of the models' prediction convolutions (pyramid.py:291-306), per level and NCHW:
    loc_l[B,4,H,W] ~ 0.5*N(0,1) and the 4-channel max-in-out confidence conf_l[B,4,H,W].  The (neg, pos) logits after
    the max-in-out reduction have neg ~ N(0,1) and gap pos - neg ~ N(-4.5, 2) like detect_inputs(mode="random");
    the channels that lose the max sit up to 2 below it.  Returns (loc_maps, conf_maps, neg_max)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    loc_maps, conf_maps, neg_max = [], [], []
    for lvl, (fw, fh) in enumerate(feature_maps(width, height, strides)):
        loc_maps.append((rng.standard_normal((B, 4, fh, fw), dtype=np.float32) * np.float32(0.5)).astype(np.float32))
        neg = rng.standard_normal((B, fh, fw), dtype=np.float32)
        pos = (neg + rng.standard_normal((B, fh, fw), dtype=np.float32) * np.float32(2.0) - np.float32(4.5)).astype(np.float32)
        top = neg if lvl == 0 else pos                      # the logit that is a max over three channels
        three = top[:, None] - rng.uniform(0.0, 2.0, (B, 3, fh, fw)).astype(np.float32)
        win = rng.integers(0, 3, (B, fh, fw))
        np.put_along_axis(three, win[:, None], top[:, None], axis=1)
        conf = np.concatenate([three, pos[:, None]], 1) if lvl == 0 else np.concatenate([neg[:, None], three], 1)
        conf_maps.append(np.ascontiguousarray(conf, dtype=np.float32))
        neg_max.append(1 if lvl == 0 else 0)
    return loc_maps, conf_maps, neg_max

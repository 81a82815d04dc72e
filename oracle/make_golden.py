"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN PYTHON CODE (build container only).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

The reference (/root/reference, read-only) has no tests or golden vectors, so
parity is pinned by executing its unmodified `layers` / `utils.calc_performance` code on seeded
synthetic inputs (fdt_b200.synth) under torch CPU, and storing inputs-by-seed (+ sha256 digest),
or small inputs verbatim, together with the reference outputs.

Harness-side shims only (the reference is never edited):
  * torch.Tensor.cuda = identity   (match_* call .cuda() on CPU tensors, box_utils.py:139-143,198-200)
  * the tracker loop lives under `if __name__ == '__main__'` behind weights + a video, so
    iouTracke_cal.py:126-155,174-176 is restated verbatim around the reference's importable
    utils.calc_performance.calculate_iou.
"""
import os
import sys
import warnings

sys.dont_write_bytecode = True
REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

import numpy as np
import torch

torch.Tensor.cuda = lambda self, *a, **k: self      # shim (1)
torch.set_num_threads(8)

from layers import Detect, PriorBoxLayer, MultiBoxLoss                    # noqa: E402  (reference)
from layers import box_utils as ref_bu                                    # noqa: E402  (reference)
from utils.calc_performance import calculate_iou as ref_iou_np            # noqa: E402  (reference)
from utils import calc_performance as ref_cp                              # noqa: E402  (reference)

from fdt_b200 import synth                                                # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
VAR = [0.1, 0.2]


def save(name, **kw):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **kw)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KB")


def ref_priors(width, height, strides=synth.STRIDES6, boxes=synth.BOXES6, **kw):
    layer = PriorBoxLayer(width, height, stride=strides, box=boxes, **kw)
    return torch.cat([layer(i, fw, fh) for i, (fw, fh) in enumerate(synth.feature_maps(width, height, strides))], 0)


# ------------------------------------------------------------------ 1. PriorBoxLayer
def gen_priorbox():
    d = {}
    for tag, (w, h, st, bx) in {
        "640x640": (640, 640, synth.STRIDES6, synth.BOXES6),
        "1024x1024": (1024, 1024, synth.STRIDES6, synth.BOXES6),
        "640x480": (640, 480, synth.STRIDES6, synth.BOXES6),
        "640x640_5lvl": (640, 640, (4, 8, 16, 32, 64), (16, 32, 64, 128, 256)),      # pyramid_mb2_try3.py:144
        "640x640_head": (640, 640, (8, 16, 32, 64, 128, 128), (16, 32, 64, 128, 256, 512)),
    }.items():
        p = ref_priors(w, h, st, bx).numpy()
        d[tag + "_sha"] = np.array(synth.digest(p)); d[tag + "_n"] = np.array(p.shape[0])
        d[tag + "_head"] = p[:64].copy(); d[tag + "_tail"] = p[-64:].copy()
    # small config with scales and aspect ratios, stored in full
    layer = PriorBoxLayer(96, 64, stride=(8, 16), box=(16, 40), scale=(3, 2), aspect_ratios=([2, 0.5], [3]))
    d["ar_l0"] = layer(0, 12, 8).numpy(); d["ar_l1"] = layer(1, 6, 4).numpy()
    save("priorbox", **d)


# ------------------------------------------------------------------ 2. elementwise box utils
def gen_boxutils():
    rng = np.random.Generator(np.random.PCG64(101))
    pri = synth.priors_numpy(160, 160)                                   # 2138 priors
    n = pri.shape[0]
    loc = (rng.standard_normal((n, 4)) * 0.7).astype(np.float32)
    dec = ref_bu.decode(torch.from_numpy(loc), torch.from_numpy(pri), VAR).numpy()
    gt = synth.gt_boxes(n, rng)[:, :4]
    enc = ref_bu.encode(torch.from_numpy(gt), torch.from_numpy(pri), VAR).numpy()
    pf = ref_bu.point_form(torch.from_numpy(pri)).numpy()
    cs = ref_bu.center_size(torch.from_numpy(pf)).numpy()
    a = synth.gt_boxes(37, rng)[:, :4]
    inter = ref_bu.intersect(torch.from_numpy(a), torch.from_numpy(pf)).numpy()
    iou = ref_bu.calculate_iou(torch.from_numpy(a), torch.from_numpy(pf)).numpy()
    x = (rng.standard_normal((500, 2)) * 3).astype(np.float32)
    lse = ref_bu.log_sum_exp(torch.from_numpy(x)).numpy()
    a64 = rng.uniform(0, 600, (50, 2)); a64 = np.concatenate([a64, a64 + rng.uniform(5, 80, (50, 2))], 1)
    b64 = rng.uniform(0, 600, (40, 2)); b64 = np.concatenate([b64, b64 + rng.uniform(5, 80, (40, 2))], 1)
    iou64 = ref_iou_np(a64, b64)
    inter64 = ref_cp.intersect(a64, b64)
    dist64 = ref_cp.calculate_distance(a64, b64)
    # calc_pr: predictions [x1,y1,x2,y2,score], truth [x,y,w,h] (My_test.py:163)
    truth_xywh = np.concatenate([b64[:12, :2], b64[:12, 2:] - b64[:12, :2]], 1)
    pred = np.concatenate([b64[:12] + rng.normal(0, 4, (12, 4)), rng.uniform(0.3, 1, (12, 1))], 1)
    pred = np.concatenate([pred, np.concatenate([a64[:9], rng.uniform(0.3, 1, (9, 1))], 1)], 0)
    pr, tnum = ref_cp.calc_pr(pred, truth_xywh)
    np.savez_compressed(os.path.join(OUT, "calcperf.npz"), a64=a64, b64=b64, inter64=inter64, dist64=dist64, truth=truth_xywh,
                        pred=pred, pr=pr, truth_num=np.int64(tnum))
    save("boxutils", priors=pri, loc=loc, decode=dec, gt=gt, encode=enc, point_form=pf, center_size=cs,
         a=a, intersect=inter, iou=iou, x=x, lse=lse, a64=a64, b64=b64, iou64=iou64)


# ------------------------------------------------------------------ 3. nms
def gen_nms():
    d = {}
    pri = synth.priors_numpy(640, 640)
    for tag, seed, n, thr, topk in (("a", 7, 3000, 0.3, 5000), ("b", 8, 6000, 0.5, 5000), ("c", 9, 300, 0.35, 200)):
        rng = np.random.Generator(np.random.PCG64(seed))
        sel = np.sort(rng.permutation(25600 + 6400)[:n])
        loc = (rng.standard_normal((n, 4)) * 0.5).astype(np.float32)
        boxes = ref_bu.decode(torch.from_numpy(loc), torch.from_numpy(pri[sel]), VAR).numpy()
        scores = np.unique(rng.uniform(0.05, 1, 2 * n).astype(np.float32))
        scores = rng.permutation(scores)[:n].copy()
        keep, count = ref_bu.nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr, topk)
        d.update({f"{tag}_boxes": boxes, f"{tag}_scores": scores, f"{tag}_thr": np.float32(thr), f"{tag}_topk": np.int64(topk),
                  f"{tag}_keep": keep.numpy()[:count].astype(np.int32), f"{tag}_count": np.int64(count)})
        print("  nms", tag, "kept", count, "of", n)
    save("nms", **d)


# ------------------------------------------------------------------ 4. Detect
def ref_detect_aux(det, loc, conf, pri):
    """Reference Detect output + kept prior indices (recomputed with the reference's own pieces)."""
    out = det(loc, conf, pri)
    B = loc.shape[0]
    kept = -np.ones((B, det.num_classes, det.top_k), np.int64)
    counts = np.zeros((B, det.num_classes), np.int32)
    for i in range(B):
        boxes = ref_bu.decode(loc[i], pri, det.variance)
        for cl in range(1, det.num_classes):
            m = conf[i, :, cl].gt(det.conf_thresh)
            nz = m.nonzero().squeeze()
            sc = conf[i, :, cl][nz]
            if sc.dim() == 0:
                continue
            bx = boxes[m.unsqueeze(1).expand_as(boxes)].view(-1, 4)
            ids, c = ref_bu.nms(bx, sc, det.nms_thresh, min(bx.shape[0], det.nms_top_k))
            c = min(c, det.top_k)
            kept[i, cl, :c] = nz[ids[:c]].numpy(); counts[i, cl] = c
            assert torch.equal(out[i, cl, :c, 0], sc[ids[:c]])
    return out.numpy(), counts, kept


def gen_detect():
    d = {}
    # (a) config 1: one 640x640 image, N = 34,125, Detect(2,0,750,0.05,0.3)   [inputs by seed]
    pri = ref_priors(640, 640)
    assert np.array_equal(pri.numpy(), synth.priors_numpy(640, 640))
    for tag, mode, seed, B in (("cfg1", "random", 20261, 1), ("clustered", "clustered", 20262, 2)):
        loc, conf = synth.detect_inputs(B, pri.numpy(), seed, 0.05, mode)
        det = Detect(2, 0, 750, 0.05, 0.3)
        out, counts, kept = ref_detect_aux(det, torch.from_numpy(loc), torch.from_numpy(conf), pri)
        print("  detect", tag, "counts", counts[:, 1], "candidates", (conf[..., 1] > 0.05).sum(1))
        d.update({f"{tag}_seed": np.int64(seed), f"{tag}_B": np.int64(B), f"{tag}_mode": np.array(mode),
                  f"{tag}_in_sha": np.array(synth.digest(loc, conf)), f"{tag}_out": out, f"{tag}_counts": counts,
                  f"{tag}_kept": kept.astype(np.int32)})
    # (b) small: head-branch-sized prior set, production thresholds, quirk images   [inputs stored]
    pri_s = ref_priors(160, 160)
    loc, conf = synth.detect_inputs(4, pri_s.numpy(), 5, 0.3, "random")
    conf[1, :, 1] = 0.01; conf[1, 777, 1] = 0.9            # Q2: exactly one candidate -> no detection
    conf[2, :, 1] = 0.2                                     # zero candidates
    conf[..., 0] = 1 - conf[..., 1]
    det = Detect(2, 0, 750, 0.3, 0.5)                       # pyramid.py:198
    out, counts, kept = ref_detect_aux(det, torch.from_numpy(loc), torch.from_numpy(conf), pri_s)
    print("  detect small counts", counts[:, 1])
    d.update(small_loc=loc, small_conf=conf, small_priors=pri_s.numpy(), small_out=out, small_counts=counts,
             small_kept=kept.astype(np.int32))
    save("detect", **d)


def gen_detect_shapes():
    """Detect at the other shapes of BASELINE.json / the reference scripts: 1024x1024 (config 5, N = 87,360) and the tracker's
    640x480 prior set (iouTracke_cal.py:98, N = 25,600) with production thresholds.  Inputs by seed."""
    d = {}
    for tag, (w, h, mode, seed, args, B) in {
        "1024": (1024, 1024, "random", 20265, (2, 0, 750, 0.05, 0.3), 1),
        "480": (640, 480, "random", 20266, (2, 0, 750, 0.3, 0.35), 1),          # ~820 candidates
        "480c": (640, 480, "clustered", 20267, (2, 0, 750, 0.3, 0.35), 2),      # a few hundred candidates around face clusters
    }.items():
        pri = ref_priors(w, h)
        assert np.array_equal(pri.numpy(), synth.priors_numpy(w, h))
        loc, conf = synth.detect_inputs(B, pri.numpy(), seed, args[3], mode)
        det = Detect(*args)
        out, counts, kept = ref_detect_aux(det, torch.from_numpy(loc), torch.from_numpy(conf), pri)
        print("  detect", tag, "N", pri.shape[0], "counts", counts[:, 1], "candidates", (conf[..., 1] > args[3]).sum(1))
        d.update({f"{tag}_cfg": np.array([w, h, seed, B]), f"{tag}_mode": np.array(mode), f"{tag}_args": np.array(args, np.float64),
                  f"{tag}_in_sha": np.array(synth.digest(loc, conf)), f"{tag}_out": out[:, 1], f"{tag}_counts": counts,
                  f"{tag}_kept": kept.astype(np.int32)})
    save("detect_shapes", **d)


# ------------------------------------------------------------------ 5. match / MultiBoxLoss
def ref_match(bipartite, thr, truth, pri, labels):
    N = pri.shape[0]
    loc_t = torch.zeros(1, N, 4); conf_t = torch.zeros(1, N, dtype=torch.long)
    fn = ref_bu.match_ensure_max_prior if bipartite else ref_bu.match_default
    fn(thr, truth, pri, VAR, labels, loc_t, conf_t, 0)
    bti = ref_bu.calculate_iou(truth, ref_bu.point_form(pri)).max(0)[1]
    return loc_t[0].numpy(), conf_t[0].numpy(), bti.numpy()


def ref_mining(crit, loc, conf, pri, targets):
    """multibox_loss.py:65-116 restated around the reference's match_* and log_sum_exp to expose
    conf_t / loss_c / neg (forward() only returns two scalars)."""
    B, N, _ = loc.shape
    loc_t = torch.Tensor(B, N, 4); conf_t = torch.LongTensor(B, N)
    fn = ref_bu.match_ensure_max_prior if crit.bipartite else ref_bu.match_default
    for i in range(B):
        fn(crit.threshold, targets[i][:, :-1], pri, crit.variance, targets[i][:, -1], loc_t, conf_t, i)
    pos = conf_t > 0
    bc = conf.view(-1, crit.num_classes)
    loss_c = ref_bu.log_sum_exp(bc) - bc.gather(1, conf_t.view(-1, 1))
    loss_c = loss_c.view(B, -1); loss_c[pos] = 0
    _, li = loss_c.sort(1, descending=True); _, rank = li.sort(1)
    num_neg = torch.clamp(crit.negpos_ratio * pos.long().sum(1, keepdim=True), max=pos.size(1) - 1)
    neg = rank < num_neg.expand_as(rank)
    # tie audit: the selected set is order-independent iff the boundary value is not duplicated
    for i in range(B):
        k = int(num_neg[i])
        if 0 < k < N:
            srt = loss_c[i].sort(descending=True)[0]
            assert srt[k - 1] != srt[k], "mining boundary tie in golden input; change the seed"
    return loc_t.numpy(), conf_t.numpy(), loss_c.numpy(), neg.numpy()


def gen_multibox():
    d = {}
    pri_s = ref_priors(160, 160)
    rng = np.random.Generator(np.random.PCG64(303))
    for tag, G in (("g1", 1), ("g7", 7), ("g60", 60)):
        gt = synth.gt_boxes(G, rng)
        for bip in (0, 1):
            lt, ct, bti = ref_match(bip, 0.35, torch.from_numpy(gt[:, :4]), pri_s, torch.from_numpy(gt[:, 4]))
            d.update({f"{tag}_gt": gt, f"{tag}_b{bip}_loc_t": lt, f"{tag}_b{bip}_conf_t": ct.astype(np.int8),
                      f"{tag}_b{bip}_bti": bti.astype(np.int16)})
    d["small_priors"] = pri_s.numpy()
    # full forward, small
    loc, conf, targets = synth.multibox_inputs(4, pri_s.numpy(), 404, 1, 20)
    for bip in (0, 1):
        crit = MultiBoxLoss(2, 0.35, True, 0, True, 3, 0.35, False, bipartite=bool(bip), use_gpu=False)
        tl = [torch.from_numpy(t) for t in targets]
        ll, lc = crit((torch.from_numpy(loc), torch.from_numpy(conf), pri_s), tl)
        lt, ct, lca, neg = ref_mining(crit, torch.from_numpy(loc), torch.from_numpy(conf), pri_s, tl)
        d.update({f"fwd_b{bip}_loss": np.array([float(ll), float(lc)], np.float64), f"fwd_b{bip}_conf_t": ct.astype(np.int8),
                  f"fwd_b{bip}_neg": np.packbits(neg), f"fwd_b{bip}_loss_c_all": lca, f"fwd_b{bip}_loc_t": lt})
        print("  multibox small bip", bip, float(ll), float(lc), "pos", int((ct > 0).sum()), "neg", int(neg.sum()))
    d["fwd_seed"] = np.int64(404); d["fwd_in_sha"] = np.array(synth.digest(loc, conf, *targets))
    # production size: B=2, N=34,125, G in [1,200]   [inputs by seed]
    pri = ref_priors(640, 640)
    loc, conf, targets = synth.multibox_inputs(2, pri.numpy(), 505, 1, 200)
    crit = MultiBoxLoss(2, 0.35, True, 0, True, 3, 0.35, False, bipartite=False, use_gpu=False)   # MyTrain_repo.py:105-114
    tl = [torch.from_numpy(t) for t in targets]
    ll, lc = crit((torch.from_numpy(loc), torch.from_numpy(conf), pri), tl)
    lt, ct, lca, neg = ref_mining(crit, torch.from_numpy(loc), torch.from_numpy(conf), pri, tl)
    bti = np.stack([ref_bu.calculate_iou(t[:, :4], ref_bu.point_form(pri)).max(0)[1].numpy() for t in tl])
    d.update(big_seed=np.int64(505), big_in_sha=np.array(synth.digest(loc, conf, *targets)),
             big_loss=np.array([float(ll), float(lc)], np.float64), big_conf_t=ct.astype(np.int8),
             big_bti=bti.astype(np.int16), big_neg=np.packbits(neg), big_G=np.array([t.shape[0] for t in targets]))
    print("  multibox big", float(ll), float(lc), "pos", int((ct > 0).sum()), "neg", int(neg.sum()))
    save("multibox", **d)


# ------------------------------------------------------------------ 6. tracker
def ref_tracker(frames, sigma_iou=0.4, sigma_h=0.6, t_min=5, use_iou=True, sigma_dis=8):
    """VERBATIM restatement of iouTracke_cal.py:126-155 (loop body, both use_iou branches) and :174-176 (flush)
    around the reference's utils.calc_performance.calculate_iou / calculate_distance."""
    frame_num = 0
    tracks_active = []
    tracks_finished = []
    for det0 in frames:
        frame_num += 1
        dets = det0.tolist()
        updated_tracks = []
        for track in tracks_active:
            if len(dets) > 0:
                if use_iou:
                    iou = ref_iou_np(np.array(dets)[:, :4], np.array([track['bboxes'][-1]]))
                    best_match = iou.argmax()
                    matched = iou[best_match] > sigma_iou
                else:
                    iou = ref_cp.calculate_distance(np.array(dets)[:, :4], np.array([track['bboxes'][-1]]))
                    best_match = iou.argmin()
                    matched = iou[best_match] < sigma_dis
                if matched:
                    track['bboxes'].append(dets[best_match][:4])
                    track['max_score'] = max(track['max_score'], dets[best_match][4])
                    updated_tracks.append(track)
                    del dets[best_match]
                else:
                    if track['max_score'] > sigma_h and len(track['bboxes']) > t_min:
                        tracks_finished.append(track)
        new_tracks = [{'bboxes': [det[:4], ], 'max_score': det[4], 'start_frame': frame_num} for det in dets]
        tracks_active = updated_tracks + new_tracks
    tracks_finished += [track for track in tracks_active
                        if track['max_score'] > sigma_h and len(track['bboxes']) >= t_min]
    return tracks_finished


def gen_tracker():
    d = {}
    for tag, kw in (("a", dict(F=400, seed=11, d_lo=1, d_hi=40, n_objects=40, empty_every=57)),
                    ("b", dict(F=120, seed=12, d_lo=1, d_hi=300, n_objects=300, empty_every=50, sigma=4.0)),
                    ("c", dict(F=300, seed=13, d_lo=1, d_hi=60, n_objects=60, empty_every=41)),            # use_iou = False
                    ("d", dict(F=100, seed=14, d_lo=1, d_hi=250, n_objects=250, empty_every=33, sigma=3.0))):  # use_iou = False, crowded
        frames = synth.tracker_frames(**kw)
        if tag in ("a", "c"):               # two consecutive empty frames: dummy-vs-dummy IoU is 0/0 = NaN (distance: 0)
            frames[200] = np.array([[0, 0, 0, 0, 0.4]]); frames[201] = np.array([[0, 0, 0, 0, 0.4]])
        tr = ref_tracker(frames, use_iou=tag in ("a", "b"))
        bb = np.array([b for t in tr for b in t['bboxes']], np.float64).reshape(-1, 4)
        d.update({f"{tag}_kw": np.array(repr(kw)), f"{tag}_in_sha": np.array(synth.digest(*frames)),
                  f"{tag}_len": np.array([len(t['bboxes']) for t in tr], np.int64),
                  f"{tag}_start": np.array([t['start_frame'] for t in tr], np.int64),
                  f"{tag}_max": np.array([t['max_score'] for t in tr], np.float64), f"{tag}_bboxes": bb})
        print("  tracker", tag, "tracks", len(tr), "boxes", bb.shape[0])
    save("tracker", **d)


def ref_detect_face_readout(detections, width, height, shrink):
    """iouTracke_cal.py:55-84 verbatim (detect_face after the network call); `detections` = y.data of ONE image."""
    scale = torch.Tensor([width, height, width, height])
    boxes = []
    scores = []
    for i in range(detections.size(1)):
        j = 0
        while detections[0, i, j, 0] >= 0.4:
            score_ = detections[0, i, j, 0]
            pt = (detections[0, i, j, 1:] * scale).cpu().numpy()
            boxes.append([pt[0], pt[1], pt[2], pt[3]])
            scores.append(score_)
            j += 1
            if j >= detections.size(2):
                break
    det_conf = np.array(scores)
    boxes = np.array(boxes)
    if boxes.shape[0] == 0:
        return np.array([[0, 0, 0, 0, 0.4]])
    det_xmin = boxes[:, 0] / shrink
    det_ymin = boxes[:, 1] / shrink
    det_xmax = boxes[:, 2] / shrink
    det_ymax = boxes[:, 3] / shrink
    det = np.column_stack((det_xmin, det_ymin, det_xmax, det_ymax, det_conf))
    keep_index = np.where(det[:, 4] >= 0)[0]
    det = det[keep_index, :]
    return det


def gen_frames():
    """Detect -> tracker chain (iouTracke_cal.py:55-84 feeding :126-155): the read-out of every frame and the tracks."""
    d = {}
    for tag, (F, top_k, seed, w, h, shrink) in {"a": (90, 24, 71, 640.0, 480.0, 1), "b": (40, 16, 72, 1280.0, 720.0, 0.5)}.items():
        det = synth.clip_detections(F, top_k, seed)
        frames = [ref_detect_face_readout(torch.from_numpy(det[f:f + 1]), w, h, shrink) for f in range(F)]
        # (rows are float32 -- numpy keeps the dtype of the float32 scalars through np.array and `/ shrink` -- the dummy row is float64)
        tr = ref_tracker(frames)
        d.update({f"{tag}_cfg": np.array([F, top_k, seed, w, h, shrink], np.float64), f"{tag}_in_sha": np.array(synth.digest(det)),
                  f"{tag}_n": np.array([fr.shape[0] for fr in frames], np.int64), f"{tag}_dets": np.concatenate(frames, 0),
                  f"{tag}_len": np.array([len(t['bboxes']) for t in tr], np.int64),
                  f"{tag}_start": np.array([t['start_frame'] for t in tr], np.int64),
                  f"{tag}_max": np.array([t['max_score'] for t in tr], np.float64),
                  f"{tag}_bboxes": np.array([b for t in tr for b in t['bboxes']], np.float64).reshape(-1, 4)})
        print("  frames", tag, "rows", int(sum(fr.shape[0] for fr in frames)), "dummy frames", int(sum(fr.shape[0] == 1 and fr[0, 4] == 0.4 and fr[0, 2] == 0 for fr in frames)), "tracks", len(tr))
    save("frames", **d)


# ------------------------------------------------------------------ 7. head post-processing
def ref_heads(loc_maps, conf_maps, softmax):
    """pyramid.py:291-309 and :331-332 restated verbatim (the model forward cannot run here: no weights, and it calls
    time.clock(), gone in Python 3.12); `tmp_conf` / `l(x)` are the prediction-convolution outputs."""
    loc, conf = [], []
    for idx, (lx, tmp_conf) in enumerate(zip(loc_maps, conf_maps)):
        if idx == 0:
            a, b, c, pos_conf = tmp_conf.chunk(4, 1)
            neg_conf = torch.cat([a, b, c], 1)
            max_conf, _ = neg_conf.max(1)
            max_conf = max_conf.view_as(pos_conf)
            conf.append(torch.cat([max_conf, pos_conf], 1).permute(0, 2, 3, 1).contiguous())
        else:
            neg_conf, a, b, c = tmp_conf.chunk(4, 1)
            pos_conf = torch.cat([a, b, c], 1)
            max_conf, _ = pos_conf.max(1)
            max_conf = max_conf.view_as(neg_conf)
            conf.append(torch.cat([neg_conf, max_conf], 1).permute(0, 2, 3, 1).contiguous())
        loc.append(lx.permute(0, 2, 3, 1).contiguous())
    loc = torch.cat([o.view(o.size(0), -1) for o in loc], 1)
    conf = torch.cat([o.view(o.size(0), -1) for o in conf], 1)
    loc = loc.view(loc.size(0), -1, 4)
    conf = conf.view(conf.size(0), -1, 2)
    if softmax:
        conf = torch.nn.Softmax(dim=-1)(conf)
    return loc, conf


def gen_heads():
    d = {}
    for tag, (B, w, h, seed, strides, boxes) in {
        "a": (2, 160, 128, 31, synth.STRIDES6, synth.BOXES6),
        "b": (2, 256, 256, 32, synth.STRIDES6[:5], synth.BOXES6[:5]),          # 5-level variant (pyramid_mb2_try3.py:144)
    }.items():
        loc_maps, conf_maps, neg_max = synth.head_maps(B, w, h, seed, strides)
        if tag == "a":
            conf_maps[1][0, 2, 3, 4] = np.nan                                  # torch.max propagates NaN
            conf_maps[0][1, 0, 5, 6] = np.inf
        tl, tc = [torch.from_numpy(m) for m in loc_maps], [torch.from_numpy(m) for m in conf_maps]
        loc, conf = ref_heads(tl, tc, True)
        _, raw = ref_heads(tl, tc, False)
        pri = ref_priors(w, h, strides, boxes)
        det = Detect(2, 0, 200, 0.05, 0.3)
        out, counts, kept = ref_detect_aux(det, loc, conf, pri)
        d.update({f"{tag}_cfg": np.array([B, w, h, seed, len(strides)]), f"{tag}_in_sha": np.array(synth.digest(*loc_maps, *conf_maps)),
                  f"{tag}_loc_sha": np.array(synth.digest(loc.numpy())), f"{tag}_raw_sha": np.array(synth.digest(raw.numpy())),
                  f"{tag}_conf": conf.numpy(),
                  f"{tag}_out": out, f"{tag}_counts": counts, f"{tag}_kept": kept.astype(np.int32)})
        print("  heads", tag, "N", loc.shape[1], "candidates", int((conf[..., 1] > 0.05).sum()))
    save("heads", **d)


# ------------------------------------------------------------------ 8. sibling decode + NMS implementations
def sibling_dets(n, seed, size=480.0, clusters=25):
    """n pixel boxes [x1,y1,x2,y2,score] fp32 around `clusters` centres (heavy overlap), unique scores."""
    rng = np.random.Generator(np.random.PCG64(seed))
    c = rng.uniform(40, size - 40, (clusters, 2))
    k = rng.integers(0, clusters, n)
    side = rng.uniform(12, 90, n)
    ctr = c[k] + rng.normal(0, 6, (n, 2))
    w = side * rng.uniform(0.8, 1.25, n); h = side * rng.uniform(0.8, 1.25, n)
    sc = rng.permutation(n).astype(np.float64) / n * 0.6 + 0.4
    d = np.stack([ctr[:, 0] - w / 2, ctr[:, 1] - h / 2, ctr[:, 0] + w / 2, ctr[:, 1] + h / 2, sc], 1).astype(np.float32)
    return d


def gen_siblings():
    sys.path.insert(0, os.path.join(REF, "FACEBOX"))
    sys.path.insert(0, os.path.join(REF, "MTCNN"))
    from encoderl import DataEncoder                                       # noqa: E402  (reference, FACEBOX/encoderl.py)
    from mtcnn.core import utils as mt_utils, nms as mt_nms                # noqa: E402  (reference, MTCNN/mtcnn/core)
    d = {}
    enc = DataEncoder()
    db = enc.default_boxes_np
    d.update(fb_default_sha=np.array(synth.digest(db)), fb_default_head=db[:40].copy(), fb_default_tail=db[-10:].copy())
    # FaceBoxes decode_np (encoderl.py:308-325): 21,824 default boxes, conf_thres 0.35, nms_np Union 0.5
    rng = np.random.Generator(np.random.PCG64(51))
    N = db.shape[0]
    loc = (rng.standard_normal((N, 4)) * 0.5).astype(np.float32)
    s1 = 1.0 / (1.0 + np.exp(-(rng.standard_normal(N) * 2.0 - 4.0)))
    s1 = synth._uniquify_candidates(s1.astype(np.float32)[None], 0.35)[0]
    conf = np.stack([1 - s1, s1], 1).astype(np.float32)
    boxes_k, scores_k = enc.decode_np(torch.from_numpy(loc), torch.from_numpy(conf))
    score = conf[:, 1]; ids = np.where(score > 0.35)[0]                    # :314-315 restated to recover the kept indices
    cxcy = loc[ids, :2] * 0.1 * db[ids, 2:] + db[ids, :2]
    wh = np.exp(loc[ids, 2:] * 0.2) * db[ids, 2:]
    bx = np.hstack([cxcy - wh / 2, cxcy + wh / 2])
    keep = np.array(enc.nms_np(bx, score[ids]), dtype=np.int64)
    assert np.array_equal(bx[keep], boxes_k) and np.array_equal(score[ids][keep], scores_k)
    d.update(fb_in_sha=np.array(synth.digest(loc, conf)), fb_kept=ids[keep].astype(np.int32), fb_boxes=boxes_k.astype(np.float32),
             fb_scores=scores_k.astype(np.float32), fb_ncand=np.int64(ids.size))
    print("  faceboxes decode_np: candidates", ids.size, "kept", keep.size)
    # FaceBoxes encode (encoderl.py:158-215): 6 faces, two of them sharing their best default box
    gtb = np.array([[0.10, 0.12, 0.22, 0.30], [0.40, 0.40, 0.47, 0.49], [0.401, 0.401, 0.471, 0.491], [0.6, 0.2, 0.95, 0.7],
                    [0.05, 0.7, 0.09, 0.76], [0.30, 0.05, 0.33, 0.085]], dtype=np.float32)
    # enc.encode itself always raises here: `if inf_flag.long().sum() is not 0` (:196) compares a tensor with `is`, true for
    # every input since torch 0.4, and the handler names an undefined `inf_error`.  Lines :171-193 and :203-206 verbatim:
    def ref_encode(boxes, classes, threshold=0.35):
        default_boxes = enc.default_boxes
        num_obj = boxes.size(0)
        iou = enc.iou(boxes, torch.cat([default_boxes[:, :2] - default_boxes[:, 2:] / 2,
                                        default_boxes[:, :2] + default_boxes[:, 2:] / 2], 1))
        max_iou, max_iou_index = iou.max(1)
        iou, max_index = iou.max(0)
        max_index.squeeze_(0)
        iou.squeeze_(0)
        max_index[max_iou_index] = torch.LongTensor(range(num_obj))
        boxes = boxes[max_index]
        variances = [0.1, 0.2]
        cxcy = (boxes[:, :2] + boxes[:, 2:]) / 2 - default_boxes[:, :2]
        cxcy /= variances[0] * default_boxes[:, 2:]
        wh = (boxes[:, 2:] - boxes[:, :2]) / default_boxes[:, 2:]
        wh = torch.log(wh) / variances[1]
        loc = torch.cat([cxcy, wh], 1)
        conf = classes[max_index]
        conf[iou < threshold] = 0
        conf[max_iou_index] = 1
        return loc, conf
    eloc, econf = ref_encode(torch.from_numpy(gtb), torch.ones(gtb.shape[0], dtype=torch.long))
    d.update(enc_gt=gtb, enc_loc=eloc.numpy(), enc_conf=econf.numpy().astype(np.int32))
    print("  faceboxes encode: positives", int((econf > 0).sum()))
    # NMS variants on the same detections
    dets = sibling_dets(600, 52)
    dets[7, :4] = dets[3, :4]                                              # identical boxes: overlap exactly 1
    dets[11, 2] = dets[11, 0]                                              # zero width
    d["dets"] = dets
    for tag, fn in (("mt_union_06", lambda: mt_utils.nms(dets, 0.6, "Union")),              # detect.py:326, :431
                    ("mt_min_04", lambda: mt_utils.nms(dets, 0.4, "Minimum")),             # detect.py:314
                    ("mt_min_05", lambda: mt_utils.nms(dets, 0.5, "Minimum")),             # detect.py:579
                    ("fb_np_union_05", lambda: enc.nms_np(dets[:, :4], dets[:, 4], 0.5)),
                    ("fb_np_min_03", lambda: enc.nms_np(dets[:, :4], dets[:, 4], 0.3, "Minimum")),
                    ("mt_plus1_union_05", lambda: mt_nms.torch_nms(dets, 0.5, "Union")),
                    ("mt_plus1_min_07", lambda: mt_nms.torch_nms(dets, 0.7, "Minimum")),
                    ("fb_torch_05", lambda: enc.nms(torch.from_numpy(dets[:, :4].copy()), torch.from_numpy(dets[:, 4].copy()), 0.5).numpy()),
                    ("fb_torch_1", lambda: enc.nms(torch.from_numpy(dets[:, :4].copy()), torch.from_numpy(dets[:, 4].copy()), 1.0).numpy())):
        k = np.asarray(fn(), dtype=np.int64).reshape(-1)
        d[tag] = k.astype(np.int32)
        print("  ", tag, "kept", k.size)
    # float64 dets, as MTCNN's pipeline passes them (core/detect.py:314, 326): boxes whose overlap sits within fp32 rounding of the
    # threshold decide differently in fp32 and float64, so these pin the float64 path
    d64 = sibling_dets(500, 54).astype(np.float64)
    rng64 = np.random.Generator(np.random.PCG64(55))
    d64[:, :4] += rng64.uniform(-1e-6, 1e-6, (500, 4))                   # not representable in fp32
    d64[:, 4] = rng64.permutation(500) / 500.0 * 0.6 + 0.4 + rng64.uniform(0, 1e-9, 500)
    d["dets64"] = d64
    d["f64_mt_union_06"] = np.asarray(mt_utils.nms(d64, 0.6, "Union"), np.int32)
    d["f64_mt_min_04"] = np.asarray(mt_utils.nms(d64, 0.4, "Minimum"), np.int32)
    d["f64_plus1_union_05"] = np.asarray(mt_nms.torch_nms(d64, 0.5, "Union"), np.int32)
    d["f64_fb_np_union_05"] = np.asarray(enc.nms_np(d64[:, :4], d64[:, 4], 0.5), np.int32)
    print("   f64 kept", d["f64_mt_union_06"].size, d["f64_mt_min_04"].size, d["f64_plus1_union_05"].size, d["f64_fb_np_union_05"].size)
    # NaN coordinate: numpy / torch min-max propagate it
    dn = sibling_dets(64, 53); dn[5, 1] = np.nan
    d["dets_nan"] = dn
    d["nan_mt_union_05"] = np.asarray(mt_utils.nms(dn, 0.5, "Union"), np.int32)
    d["nan_mt_min_05"] = np.asarray(mt_utils.nms(dn, 0.5, "Minimum"), np.int32)
    d["nan_plus1_union_05"] = np.asarray(mt_nms.torch_nms(dn, 0.5, "Union"), np.int32)
    save("siblings", **d)


# ------------------------------------------------------------------ 9. on-disk formats
def gen_formats():
    from data.widerface import AnnotationTransform as RefAT                      # noqa: E402  (reference)
    lines = ["a/b/img0.jpg 3 10 20 30 40 5 6 0 9 100 120 -20 35",
             "c.jpg 2 7 8 9 -10 1 2 3 4",
             "d.jpg 0",
             open(os.path.join(REF, "image_and_anno/anno/gen_anno_file_val")).readline().strip()]
    d = {"lines": np.array(lines)}
    for i, ln in enumerate(lines):
        f = ln.strip().split()
        d[f"at_{i}"] = np.array(RefAT()(f[1:], 640, 480), dtype=np.float64).reshape(-1, 5)
        num = int(f[1]); tgt = list(f[1:]); del tgt[0]                           # utils/data_collector.py:46-51
        d[f"px_{i}"] = np.array(tgt).astype(np.int32).reshape(num, 4)
    # PR data: My_test.py:105, :161, :166-171 and draw_pr_roc.py:5-19, :27-34 restated (both are scripts, not importable)
    rng = np.random.Generator(np.random.PCG64(61))
    tf_conf = np.array([[], []]); truth_num = 0
    for _ in range(5):
        n = int(rng.integers(1, 40))
        tfi = np.vstack(((rng.uniform(size=n) > 0.4).astype(np.int32), rng.uniform(0.0, 1.0, n)))
        tf_conf = np.hstack((tf_conf, tfi)); truth_num += int(rng.integers(1, 30))
    d["pr_tf_conf"] = tf_conf; d["pr_truth_num"] = np.int64(truth_num)
    srt = tf_conf[:, np.argsort(tf_conf[1, :])[::-1]]
    data = np.hstack((srt, [[0], [truth_num]]))
    d["pr_file"] = data
    _, M = data[:, :-1].shape
    tp, fp = np.zeros(M), np.zeros(M)
    for i in range(1, M + 1):
        tp[i - 1] = np.count_nonzero(data[0, :i]); fp[i - 1] = i - tp[i - 1]
    d["pr_tp"] = tp; d["pr_fp"] = fp; d["pr_recall"] = tp / data[1, -1]; d["pr_precision"] = tp / (tp + fp)
    save("formats", **d)


if __name__ == "__main__":
    which = sys.argv[1:] or ["priorbox", "boxutils", "nms", "detect", "multibox", "tracker", "heads", "siblings", "formats", "detect_shapes", "frames"]
    for w in which:
        print("==", w)
        globals()["gen_" + w]()

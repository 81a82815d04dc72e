"""Pins the CPU oracle (oracle/fdt_oracle.c) against outputs of the reference's own Python code
(tests/golden/*.npz, produced by oracle/make_golden.py).  Runs without a GPU."""
import numpy as np
import pytest

from fdt_b200.synth import clip_detections

from fdt_b200 import synth
from oracle import oracle as orc

VAR = (0.1, 0.2)
RTOL = 1e-5      # coordinates that went through exp/log (north_star tolerance)


def close(a, b, rtol=RTOL, atol=1e-7):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def orc_priors(width, height, strides=synth.STRIDES6, boxes=synth.BOXES6):
    layer = orc.PriorBoxLayer(width, height, stride=strides, box=boxes)
    return np.concatenate([layer(i, fw, fh) for i, (fw, fh) in enumerate(synth.feature_maps(width, height, strides))], 0)


@pytest.mark.parametrize("tag,w,h,st,bx", [
    ("640x640", 640, 640, synth.STRIDES6, synth.BOXES6),
    ("1024x1024", 1024, 1024, synth.STRIDES6, synth.BOXES6),
    ("640x480", 640, 480, synth.STRIDES6, synth.BOXES6),
    ("640x640_5lvl", 640, 640, (4, 8, 16, 32, 64), (16, 32, 64, 128, 256)),
    ("640x640_head", 640, 640, (8, 16, 32, 64, 128, 128), (16, 32, 64, 128, 256, 512)),
])
def test_priorbox_bit_exact(golden, tag, w, h, st, bx):
    g = golden("priorbox")
    p = orc_priors(w, h, st, bx)
    assert p.shape[0] == int(g[tag + "_n"])
    assert np.array_equal(p[:64], g[tag + "_head"]) and np.array_equal(p[-64:], g[tag + "_tail"])
    assert synth.digest(p) == str(g[tag + "_sha"])


def test_priorbox_scales_and_aspect_ratios(golden):
    g = golden("priorbox")
    layer = orc.PriorBoxLayer(96, 64, stride=(8, 16), box=(16, 40), scale=(3, 2), aspect_ratios=([2, 0.5], [3]))
    assert np.array_equal(layer(0, 12, 8), g["ar_l0"])
    assert np.array_equal(layer(1, 6, 4), g["ar_l1"])


def test_synth_priors_match_reference(golden):
    g = golden("priorbox")
    assert synth.digest(synth.priors_numpy(640, 640)) == str(g["640x640_sha"])
    assert synth.digest(synth.priors_numpy(1024, 1024)) == str(g["1024x1024_sha"])


def test_elementwise_box_utils(golden):
    g = golden("boxutils")
    assert np.array_equal(orc.point_form(g["priors"]), g["point_form"])
    assert np.array_equal(orc.center_size(g["point_form"]), g["center_size"])
    assert np.array_equal(orc.intersect(g["a"], g["point_form"]), g["intersect"])
    assert np.array_equal(orc.calculate_iou(g["a"], g["point_form"]), g["iou"])
    close(orc.decode(g["loc"], g["priors"], VAR), g["decode"])
    close(orc.encode(g["gt"], g["priors"], VAR), g["encode"])
    close(orc.log_sum_exp(g["x"]), g["lse"])
    assert np.array_equal(orc.calculate_iou_f64(g["a64"], g["b64"]), g["iou64"])


def test_decode_exact_part(golden):
    """x1y1/x2y2 only differ from the reference through exp(): cx,cy path is bit-exact."""
    g = golden("boxutils")
    d = orc.decode(g["loc"], g["priors"], VAR)
    frac_equal = np.mean(d == g["decode"])
    assert frac_equal > 0.97          # SLEEF exp differs from correctly-rounded exp in ~1 % of elements


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_nms_bit_exact(golden, tag):
    g = golden("nms")
    keep, count = orc.nms(g[f"{tag}_boxes"], g[f"{tag}_scores"], float(g[f"{tag}_thr"]), int(g[f"{tag}_topk"]))
    assert count == int(g[f"{tag}_count"])
    assert np.array_equal(keep[:count], g[f"{tag}_keep"])
    assert not keep[count:].any()


def check_detect(out, counts, kept, g, tag):
    assert np.array_equal(counts, g[f"{tag}_counts"])
    assert np.array_equal(kept, g[f"{tag}_kept"])
    ref = g[f"{tag}_out"]
    assert np.array_equal(out[..., 0], ref[..., 0])            # scores are copied, not computed
    close(out[..., 1:], ref[..., 1:])
    assert not out[:, 0].any()


@pytest.mark.parametrize("tag", ["cfg1", "clustered"])
def test_detect_production_size(golden, tag):
    g = golden("detect")
    pri = synth.priors_numpy(640, 640)
    loc, conf = synth.detect_inputs(int(g[f"{tag}_B"]), pri, int(g[f"{tag}_seed"]), 0.05, str(g[f"{tag}_mode"]))
    assert synth.digest(loc, conf) == str(g[f"{tag}_in_sha"]), "synthetic generator drifted"
    det = orc.Detect(2, 0, 750, 0.05, 0.3)
    for early in (False, True):
        det.early_exit = early
        out, counts, kept = det(loc, conf, pri, return_aux=True)
        check_detect(out, counts, kept, g, tag)


def test_detect_small_with_quirks(golden):
    g = golden("detect")
    det = orc.Detect(2, 0, 750, 0.3, 0.5)
    out, counts, kept = det(g["small_loc"], g["small_conf"], g["small_priors"], return_aux=True)
    check_detect(out, counts, kept, g, "small")
    assert counts[1, 1] == 0 and counts[2, 1] == 0      # Q2 single candidate, and zero candidates


def test_detect_rejects_nonpositive_nms_thresh():
    with pytest.raises(ValueError):
        orc.Detect(2, 0, 750, 0.05, 0.0)


@pytest.mark.parametrize("tag", ["g1", "g7", "g60"])
@pytest.mark.parametrize("bip", [0, 1])
def test_match(golden, tag, bip):
    g = golden("multibox")
    gt = g[f"{tag}_gt"]
    loc_t, conf_t, bti, _ = orc.match(bip, 0.35, gt[:, :4], g["small_priors"], VAR, gt[:, 4])
    assert np.array_equal(conf_t, g[f"{tag}_b{bip}_conf_t"])
    if not bip:
        assert np.array_equal(bti, g[f"{tag}_b{bip}_bti"])
    close(loc_t, g[f"{tag}_b{bip}_loc_t"], atol=1e-6)


def test_match_zero_gt_raises():
    with pytest.raises(IndexError):
        orc.match(0, 0.35, np.zeros((0, 4), np.float32), synth.priors_numpy(160, 160), VAR, np.zeros(0, np.float32))


@pytest.mark.parametrize("bip", [0, 1])
def test_multibox_forward_small(golden, bip):
    g = golden("multibox")
    pri = g["small_priors"]
    loc, conf, targets = synth.multibox_inputs(4, pri, int(g["fwd_seed"]), 1, 20)
    assert synth.digest(loc, conf, *targets) == str(g["fwd_in_sha"])
    r = orc.multibox_loss(loc, conf, pri, targets, 0.35, 3, bool(bip), VAR)
    assert np.array_equal(r["conf_t"], g[f"fwd_b{bip}_conf_t"])
    close(r["loc_t"], g[f"fwd_b{bip}_loc_t"], atol=1e-6)
    close(r["loss_c_all"], g[f"fwd_b{bip}_loss_c_all"], atol=1e-6)
    neg_ref = np.unpackbits(g[f"fwd_b{bip}_neg"])[:r["neg"].size].reshape(r["neg"].shape).astype(bool)
    assert np.array_equal(r["neg"], neg_ref)
    close([r["loss_l"], r["loss_c"]], g[f"fwd_b{bip}_loss"])
    # mining on the reference's own loss_c is bit-exact
    neg2 = orc.hard_negative_mine(g[f"fwd_b{bip}_loss_c_all"], r["conf_t"] > 0, 3)
    assert np.array_equal(neg2, neg_ref)


def test_multibox_forward_production_size(golden):
    g = golden("multibox")
    pri = synth.priors_numpy(640, 640)
    loc, conf, targets = synth.multibox_inputs(2, pri, int(g["big_seed"]), 1, 200)
    assert synth.digest(loc, conf, *targets) == str(g["big_in_sha"])
    r = orc.multibox_loss(loc, conf, pri, targets, 0.35, 3, False, VAR)
    assert np.array_equal(r["conf_t"], g["big_conf_t"])
    neg_ref = np.unpackbits(g["big_neg"])[:r["neg"].size].reshape(r["neg"].shape).astype(bool)
    assert np.array_equal(r["neg"], neg_ref)
    close([r["loss_l"], r["loss_c"]], g["big_loss"])
    for b, t in enumerate(targets):
        _, _, bti, _ = orc.match(0, 0.35, t[:, :4], pri, VAR, t[:, 4])
        assert np.array_equal(bti, g["big_bti"][b])


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_tracker_bit_exact(golden, tag):
    """a, b: use_iou = True; c, d: use_iou = False (calculate_distance, argmin, < sigma_dis; iouTracke_cal.py:135-138)."""
    g = golden("tracker")
    kw = eval(str(g[f"{tag}_kw"]))
    frames = synth.tracker_frames(**kw)
    if tag in ("a", "c"):
        frames[200] = np.array([[0, 0, 0, 0, 0.4]]); frames[201] = np.array([[0, 0, 0, 0, 0.4]])
    assert synth.digest(*frames) == str(g[f"{tag}_in_sha"])
    tr = orc.iou_track(frames, use_iou=tag in ("a", "b"))
    assert [len(t["bboxes"]) for t in tr] == g[f"{tag}_len"].tolist()
    assert [t["start_frame"] for t in tr] == g[f"{tag}_start"].tolist()
    assert np.array_equal(np.array([t["max_score"] for t in tr]), g[f"{tag}_max"])
    bb = np.array([b for t in tr for b in t["bboxes"]], np.float64).reshape(-1, 4)
    assert np.array_equal(bb, g[f"{tag}_bboxes"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_detections_to_frames_and_tracks(golden, tag):
    """Detect -> tracker chain of iouTracke_cal.py:55-84 + :126-155: per-frame read-out (incl. the dummy row) and the tracks."""
    g = golden("frames")
    F, top_k, seed, w, h, shrink = g[f"{tag}_cfg"]
    det = clip_detections(int(F), int(top_k), int(seed))
    assert synth.digest(det) == str(g[f"{tag}_in_sha"])
    frames = orc.detections_to_frames(det, float(w), float(h), 0.4, float(shrink))
    assert [f.shape[0] for f in frames] == g[f"{tag}_n"].tolist()
    assert np.array_equal(np.concatenate([np.asarray(f, np.float64) for f in frames], 0), g[f"{tag}_dets"])
    tr = orc.iou_track(frames)
    assert [len(t["bboxes"]) for t in tr] == g[f"{tag}_len"].tolist() and [t["start_frame"] for t in tr] == g[f"{tag}_start"].tolist()
    assert np.array_equal(np.array([b for t in tr for b in t["bboxes"]], np.float64).reshape(-1, 4), g[f"{tag}_bboxes"])


def test_calc_performance_functions(golden):
    g = golden("calcperf")
    assert np.array_equal(orc.intersect_f64(g["a64"], g["b64"]), g["inter64"])
    np.testing.assert_allclose(orc.calculate_distance_f64(g["a64"], g["b64"]), g["dist64"], rtol=1e-14)
    pr, tnum = orc.calc_pr(g["pred"], g["truth"])
    assert tnum == int(g["truth_num"]) and np.array_equal(pr, g["pr"])
    assert pr[0].sum() >= 10                                   # the fixture really has matches and misses


def test_torch_restatement_matches_oracle():
    """oracle/torch_ref.py (the reference's python loops restated in torch, used only to time the reference's intended
    torch-CUDA deployment on the GPU box) agrees with the C oracle: same kept rows, coordinates within the exp tolerance."""
    import torch
    from oracle import torch_ref
    pri = synth.priors_numpy(160, 160)
    loc, conf = synth.detect_inputs(2, pri, 5, 0.05)
    out = torch_ref.Detect(2, 0, 750, 0.05, 0.3)(torch.from_numpy(loc), torch.from_numpy(conf), torch.from_numpy(pri)).numpy()
    ref = orc.Detect(2, 0, 750, 0.05, 0.3)(loc, conf, pri)
    assert np.array_equal(out[..., 0], ref[..., 0])
    close(out, ref)


def shapes_case(g, tag):
    w, h, seed, B = (int(v) for v in g[tag + "_cfg"])
    a = g[tag + "_args"]
    args = (int(a[0]), int(a[1]), int(a[2]), float(a[3]), float(a[4]))
    pri = synth.priors_numpy(w, h)
    loc, conf = synth.detect_inputs(B, pri, seed, args[3], str(g[tag + "_mode"]))
    assert synth.digest(loc, conf) == str(g[tag + "_in_sha"]), "synthetic generator drifted"
    return loc, conf, pri, args


@pytest.mark.parametrize("tag", ["1024", "480", "480c"])
def test_detect_other_shapes(golden, tag):
    """1024x1024 (BASELINE config 5, N = 87,360) and the tracker's 640x480 prior set with production thresholds."""
    g = golden("detect_shapes")
    loc, conf, pri, args = shapes_case(g, tag)
    out, counts, kept = orc.Detect(*args)(loc, conf, pri, return_aux=True)
    assert np.array_equal(counts, g[tag + "_counts"]) and np.array_equal(kept, g[tag + "_kept"])
    assert np.array_equal(out[:, 1, :, 0], g[tag + "_out"][..., 0])
    close(out[:, 1, :, 1:], g[tag + "_out"][..., 1:])

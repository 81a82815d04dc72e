"""GPU parity tests for the IoU tracker association and the tracker-side calculate_iou (float64)."""
import os

import numpy as np
import pytest

from fdt_b200 import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tracker():
    from fdt_b200 import tracker as T
    return T


def same_tracks(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x["start_frame"] == y["start_frame"] and x["max_score"] == y["max_score"]
        assert x["bboxes"] == y["bboxes"]                      # float64 boxes bit-exact, same append order


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
@pytest.mark.parametrize("slow", ["0", "1"])
def test_tracker_golden(tracker, golden, tag, slow, monkeypatch):
    """a, b: use_iou = True; c, d: use_iou = False (calculate_distance / argmin / < sigma_dis, iouTracke_cal.py:135-138)."""
    monkeypatch.setenv("FDT_TRACK_FORCE_SLOW", slow)          # exercise both the parallel and the sequential path
    g = golden("tracker")
    frames = synth.tracker_frames(**eval(str(g[f"{tag}_kw"])))
    if tag in ("a", "c"):
        frames[200] = np.array([[0, 0, 0, 0, 0.4]]); frames[201] = np.array([[0, 0, 0, 0, 0.4]])
    assert synth.digest(*frames) == str(g[f"{tag}_in_sha"])
    tr = tracker.iou_track(frames, use_iou=tag in ("a", "b"))
    assert [len(t["bboxes"]) for t in tr] == g[f"{tag}_len"].tolist()          # track IDs = list order
    assert [t["start_frame"] for t in tr] == g[f"{tag}_start"].tolist()
    assert np.array_equal(np.array([t["max_score"] for t in tr]), g[f"{tag}_max"])
    bb = np.array([b for t in tr for b in t["bboxes"]], np.float64).reshape(-1, 4)
    assert np.array_equal(bb, g[f"{tag}_bboxes"])


@pytest.mark.parametrize("kw", [
    dict(F=10000, seed=4040, d_lo=1, d_hi=300, n_objects=300, empty_every=1000),        # BASELINE config 4 at full length (bench.py's video)
    dict(F=2000, seed=21, d_lo=1, d_hi=300, n_objects=300, empty_every=333),            # config-4 shape, shorter
    dict(F=500, seed=22, d_lo=100, d_hi=120, n_objects=120, empty_every=0, sigma=6.0),  # crowded, many conflicts
    dict(F=300, seed=23, d_lo=1, d_hi=3, n_objects=3, empty_every=7),
    dict(F=50, seed=24, d_lo=700, d_hi=750, n_objects=750, empty_every=0),              # Detect's top_k = 750 per frame
])
def test_tracker_vs_oracle(tracker, kw):
    frames = synth.tracker_frames(**kw)
    same_tracks(tracker.iou_track(frames), orc.iou_track(frames))


@pytest.mark.parametrize("kw,sigma_dis", [
    (dict(F=1500, seed=51, d_lo=1, d_hi=300, n_objects=300, empty_every=250), 8),
    (dict(F=400, seed=52, d_lo=80, d_hi=120, n_objects=120, empty_every=0, sigma=6.0), 20),          # loose gate: many conflicts
    (dict(F=300, seed=53, d_lo=1, d_hi=4, n_objects=4, empty_every=9), 2),
])
def test_tracker_distance_mode_vs_oracle(tracker, kw, sigma_dis):
    frames = synth.tracker_frames(**kw)
    same_tracks(tracker.iou_track(frames, use_iou=False, sigma_dis=sigma_dis), orc.iou_track(frames, use_iou=False, sigma_dis=sigma_dis))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_detections_to_frames_chain_on_device(tracker, golden, tag):
    """Detect output of a clip -> packed float64 frames (+ dummy rows) -> tracks without leaving the device, against the reference's
    detect_face read-out and tracker loop (iouTracke_cal.py:55-84, :126-155; fixture made by oracle/make_golden.py::gen_frames)."""
    import torch
    g = golden("frames")
    F, top_k, seed, w, h, shrink = g[f"{tag}_cfg"]
    det = synth.clip_detections(int(F), int(top_k), int(seed))
    assert synth.digest(det) == str(g[f"{tag}_in_sha"])
    d = torch.from_numpy(det).cuda()
    dets, off = tracker.detections_to_frames(d, float(w), float(h), 0.4, float(shrink))
    assert dets.dtype == torch.float64 and dets.is_cuda and off.is_cuda
    assert np.array_equal(np.diff(off.cpu().numpy()), g[f"{tag}_n"])
    assert np.array_equal(dets.cpu().numpy(), g[f"{tag}_dets"])
    tr = tracker.track_detections(d, float(w), float(h), 0.4, float(shrink))
    assert [len(t["bboxes"]) for t in tr] == g[f"{tag}_len"].tolist() and [t["start_frame"] for t in tr] == g[f"{tag}_start"].tolist()
    assert np.array_equal(np.array([t["max_score"] for t in tr]), g[f"{tag}_max"])
    assert np.array_equal(np.array([b for t in tr for b in t["bboxes"]], np.float64).reshape(-1, 4), g[f"{tag}_bboxes"])
    # a longer clip (more than one scan chunk of frames) and NaN / all-below-threshold planes against the oracle
    det2 = synth.clip_detections(2500, 12, 99)
    det2[7, 1, 0, 0] = np.nan; det2[8, 1, :, 0] = 0.39
    d2 = torch.from_numpy(det2).cuda()
    dets2, off2 = tracker.detections_to_frames(d2, 640.0, 480.0, 0.4, 1.0)
    ref = orc.detections_to_frames(det2, 640.0, 480.0, 0.4, 1.0)
    assert np.array_equal(np.diff(off2.cpu().numpy()), [f.shape[0] for f in ref])
    assert np.array_equal(dets2.cpu().numpy(), np.concatenate([np.asarray(f, np.float64) for f in ref], 0))
    same_tracks(tracker.track_detections(d2, 640.0, 480.0), orc.iou_track(ref))


def test_tracker_parameters_and_edge_frames(tracker):
    frames = synth.tracker_frames(F=200, seed=31, d_lo=1, d_hi=30, n_objects=30, empty_every=0)
    frames[10] = np.zeros((0, 5))                            # truly empty frame: every active track silently dropped (Q5)
    frames[50] = frames[49].copy()                           # identical frames: IoU == 1 ties resolved in order
    frames[60] = np.concatenate([frames[60], frames[60][:3]])   # duplicated detections compete for one track
    for args in ((0.4, 0.6, 5), (0.1, 0.0, 0), (0.7, 0.9, 2)):
        same_tracks(tracker.iou_track(frames, *args), orc.iou_track(frames, *args))
    assert tracker.iou_track([]) == []


def test_tracker_npy_layout_roundtrip(tracker, tmp_path):
    frames = synth.tracker_frames(F=100, seed=41, d_lo=1, d_hi=20, n_objects=20, empty_every=0)
    tr = tracker.iou_track(frames)
    tracker.save_tracks(str(tmp_path / "video8"), tr)        # np.save appends .npy (iouTracke_cal.py:177)
    loaded = np.load(str(tmp_path / "video8.npy"), allow_pickle=True).tolist()     # iouTracke_display.py:29
    assert loaded == tr and set(loaded[0]) == {"bboxes", "max_score", "start_frame"}
    assert tracker.load_tracks(str(tmp_path / "video8.npy")) == tr


def test_calc_performance_iou_f64(golden):
    from fdt_b200.utils.calc_performance import calculate_iou
    g = golden("boxutils")
    out = calculate_iou(g["a64"], g["b64"])
    assert out.dtype == np.float64 and np.array_equal(out, g["iou64"])
    track_box = np.array([g["b64"][3]])
    assert np.array_equal(calculate_iou(g["a64"], track_box), g["iou64"][:, 3:4])          # the tracker's [D,1] call shape


def test_calc_performance_intersect_distance_calc_pr(golden):
    from fdt_b200.utils import calc_performance as cp
    g = golden("calcperf")
    assert np.array_equal(cp.intersect(g["a64"], g["b64"]), g["inter64"])
    np.testing.assert_allclose(cp.calculate_distance(g["a64"], g["b64"]), g["dist64"], rtol=1e-14)
    pr, tnum = cp.calc_pr(g["pred"], g["truth"])
    assert tnum == int(g["truth_num"]) and np.array_equal(pr, g["pr"])
    pred = g["pred"].copy(); pred[3, 0] = np.nan                # NaN IoU -> max is NaN -> not matched (numpy semantics)
    assert np.array_equal(cp.calc_pr(pred, g["truth"])[0], orc.calc_pr(pred, g["truth"])[0])

"""GPU parity tests for PriorBox / box utils / nms / Detect: CUDA path (through the C ABI) vs the CPU
oracle on the same seeded inputs, and vs the reference-generated golden vectors."""
import numpy as np
import pytest
import torch

from fdt_b200 import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
VAR = (0.1, 0.2)
RTOL = 1e-5          # north_star: decoded/encoded coordinates within 1e-5 relative in fp32


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def npy(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def layers():
    import fdt_b200.layers as L
    return L


# ------------------------------------------------------------------------------- PriorBox
@pytest.mark.parametrize("w,h,st,bx", [
    (640, 640, synth.STRIDES6, synth.BOXES6), (1024, 1024, synth.STRIDES6, synth.BOXES6),
    (640, 480, synth.STRIDES6, synth.BOXES6), (640, 640, (4, 8, 16, 32, 64), (16, 32, 64, 128, 256)),
])
def test_priorbox_bit_exact(layers, w, h, st, bx):
    ours = layers.PriorBoxLayer(w, h, stride=st, box=bx)
    ref = orc.PriorBoxLayer(w, h, stride=st, box=bx)
    for i, (fw, fh) in enumerate(synth.feature_maps(w, h, st)):
        assert np.array_equal(npy(ours(i, fw, fh)), ref(i, fw, fh))


def test_priorbox_scales_and_aspect_ratios(layers, golden):
    g = golden("priorbox")
    layer = layers.PriorBoxLayer(96, 64, stride=(8, 16), box=(16, 40), scale=(3, 2), aspect_ratios=([2, 0.5], [3]))
    assert np.array_equal(npy(layer(0, 12, 8)), g["ar_l0"])
    assert np.array_equal(npy(layer(1, 6, 4)), g["ar_l1"])


def test_priorbox_640_matches_reference_digest(layers, golden):
    g = golden("priorbox")
    layer = layers.PriorBoxLayer(640, 640)
    p = torch.cat([layer(i, fw, fh) for i, (fw, fh) in enumerate(synth.feature_maps(640, 640))], 0)
    assert synth.digest(npy(p)) == str(g["640x640_sha"])


# ------------------------------------------------------------------------------- box utils
def test_elementwise_box_utils(layers, golden):
    g = golden("boxutils")
    bu = layers.box_utils
    assert np.array_equal(npy(bu.point_form(cu(g["priors"]))), g["point_form"])
    assert np.array_equal(npy(bu.center_size(cu(g["point_form"]))), g["center_size"])
    assert np.array_equal(npy(bu.intersect(cu(g["a"]), cu(g["point_form"]))), g["intersect"])
    assert np.array_equal(npy(bu.calculate_iou(cu(g["a"]), cu(g["point_form"]))), g["iou"])
    dec = npy(bu.decode(cu(g["loc"]), cu(g["priors"]), VAR))
    assert np.array_equal(dec, orc.decode(g["loc"], g["priors"], VAR))           # same exp rule -> bit-exact
    np.testing.assert_allclose(dec, g["decode"], rtol=RTOL, atol=1e-7)          # vs torch (SLEEF exp)
    enc = npy(bu.encode(cu(g["gt"]), cu(g["priors"]), VAR))
    np.testing.assert_allclose(enc, orc.encode(g["gt"], g["priors"], VAR), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(enc, g["encode"], rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(npy(bu.log_sum_exp(cu(g["x"]))), g["lse"], rtol=RTOL, atol=1e-6)


def test_box_utils_accept_cpu_tensors(layers, golden):
    g = golden("boxutils")
    out = layers.box_utils.point_form(torch.from_numpy(g["priors"]))
    assert out.device.type == "cpu" and np.array_equal(out.numpy(), g["point_form"])


# ------------------------------------------------------------------------------- nms
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_nms_golden(layers, golden, tag):
    g = golden("nms")
    thr, topk = float(g[f"{tag}_thr"]), int(g[f"{tag}_topk"])
    n = g[f"{tag}_scores"].shape[0]
    if min(n, topk) > 5000:
        pytest.skip("run-to-completion nms holds at most 5000 candidates in shared memory")
    keep, count = layers.box_utils.nms(cu(g[f"{tag}_boxes"]), cu(g[f"{tag}_scores"]), thr, topk)
    assert count == int(g[f"{tag}_count"])
    assert np.array_equal(npy(keep)[:count], g[f"{tag}_keep"])
    assert not npy(keep)[count:].any()


@pytest.mark.parametrize("n,topk,thr,seed", [(1, 200, 0.5, 0), (2, 200, 0.5, 1), (63, 200, 0.4, 2), (64, 10, 0.3, 3),
                                              (65, 200, 0.3, 4), (1000, 200, 0.5, 5), (4999, 5000, 0.3, 6),
                                              (5000, 5000, 0.45, 7), (9000, 5000, 0.3, 8), (20000, 3000, 0.6, 9)])
def test_nms_vs_oracle(layers, n, topk, thr, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    ctr = rng.uniform(0.1, 0.9, (n, 2)); wh = rng.uniform(0.02, 0.15, (n, 2))
    boxes = np.concatenate([ctr - wh / 2, ctr + wh / 2], 1).astype(np.float32)
    scores = rng.permutation(np.unique(rng.uniform(0, 1, 3 * n).astype(np.float32)))[:n].copy()
    keep, count = layers.box_utils.nms(cu(boxes), cu(scores), thr, topk)
    rk, rc = orc.nms(boxes, scores, thr, topk)
    assert count == rc and np.array_equal(npy(keep), rk)


def test_nms_tie_rule_and_degenerate_boxes(layers):
    """equal scores: higher index first; zero-area duplicates: 0/0 = NaN IoU suppresses (box_utils.py:339)."""
    boxes = np.array([[0, 0, 1, 1], [5, 5, 6, 6], [0, 0, 1, 1], [2, 2, 2, 2], [2, 2, 2, 2]], np.float32)
    scores = np.array([0.5, 0.5, 0.9, 0.3, 0.2], np.float32)
    keep, count = layers.box_utils.nms(cu(boxes), cu(scores), 0.5, 200)
    rk, rc = orc.nms(boxes, scores, 0.5, 200)
    assert count == rc and np.array_equal(npy(keep), rk)
    assert npy(keep)[:count].tolist() == [2, 1, 3]


def test_nms_empty(layers):
    keep, count = layers.box_utils.nms(torch.zeros(0, 4).cuda(), torch.zeros(0).cuda(), 0.5, 200)
    assert count == 0 and keep.numel() == 0


# ------------------------------------------------------------------------------- Detect
def run_detect(layers, loc, conf, pri, args=(2, 0, 750, 0.05, 0.3), host=False, nms_top_k=None):
    det = layers.Detect(*args)
    if nms_top_k:
        det.nms_top_k = nms_top_k
    f = (lambda a: torch.from_numpy(a)) if host else cu
    out, counts, kept = det(f(loc), f(conf), f(pri), return_aux=True)
    return npy(out), npy(counts), npy(kept)


def oracle_detect(loc, conf, pri, args=(2, 0, 750, 0.05, 0.3), nms_top_k=None):
    det = orc.Detect(*args); det.early_exit = True
    if nms_top_k:
        det.nms_top_k = nms_top_k
    return det(loc, conf, pri, return_aux=True)


def assert_same(a, b):
    for x, y, name in zip(a, b, ("out", "counts", "kept_prior")):
        assert np.array_equal(x, y), f"{name} differs from the oracle"


@pytest.mark.parametrize("tag", ["cfg1", "clustered"])
def test_detect_golden_production_size(layers, golden, tag):
    g = golden("detect")
    pri = synth.priors_numpy(640, 640)
    loc, conf = synth.detect_inputs(int(g[f"{tag}_B"]), pri, int(g[f"{tag}_seed"]), 0.05, str(g[f"{tag}_mode"]))
    assert synth.digest(loc, conf) == str(g[f"{tag}_in_sha"])
    out, counts, kept = run_detect(layers, loc, conf, pri)
    assert np.array_equal(counts, g[f"{tag}_counts"])
    assert np.array_equal(kept, g[f"{tag}_kept"])                       # kept-box indices bit-exact vs the reference
    assert np.array_equal(out[..., 0], g[f"{tag}_out"][..., 0])
    np.testing.assert_allclose(out[..., 1:], g[f"{tag}_out"][..., 1:], rtol=RTOL, atol=1e-7)
    assert_same((out, counts, kept), oracle_detect(loc, conf, pri))


def test_detect_golden_small_with_quirks(layers, golden):
    g = golden("detect")
    args = (2, 0, 750, 0.3, 0.5)
    out, counts, kept = run_detect(layers, g["small_loc"], g["small_conf"], g["small_priors"], args)
    assert np.array_equal(counts, g["small_counts"]) and np.array_equal(kept, g["small_kept"])
    assert counts[1, 1] == 0 and counts[2, 1] == 0           # single candidate (Q2) and no candidate
    np.testing.assert_allclose(out, g["small_out"], rtol=RTOL, atol=1e-7)
    assert_same((out, counts, kept), oracle_detect(g["small_loc"], g["small_conf"], g["small_priors"], args))


def test_detect_host_buffers_equal_device_path(layers, golden):
    g = golden("detect")
    args = (2, 0, 750, 0.3, 0.5)
    dev = run_detect(layers, g["small_loc"], g["small_conf"], g["small_priors"], args)
    hst = run_detect(layers, g["small_loc"], g["small_conf"], g["small_priors"], args, host=True)
    assert_same(dev, hst)
    det = layers.Detect(*args)
    o = det(torch.from_numpy(g["small_loc"]), torch.from_numpy(g["small_conf"]), torch.from_numpy(g["small_priors"]))
    assert o.device.type == "cpu" and tuple(o.shape) == (4, 2, 750, 5)


def test_detect_host_pinned_buffers_zero_copy_loc(layers):
    """pinned loc is gathered in place over PCIe (no H2D copy of loc); pageable loc is copied: same bits either way."""
    pri = synth.priors_numpy(640, 640)
    loc, conf = synth.detect_inputs(4, pri, 123, 0.05)
    det = layers.Detect(2, 0, 750, 0.05, 0.3)
    pinned = [torch.from_numpy(a).pin_memory() for a in (loc, conf, pri)]
    pageable = [torch.from_numpy(a) for a in (loc, conf, pri)]
    o1 = det(*pinned); o2 = det(*pageable)
    o3 = det(pinned[0][:, :, :].view(4, -1), pinned[1], pinned[2])
    ref = oracle_detect(loc, conf, pri)[0]
    for o in (o1, o2, o3):
        assert o.device.type == "cpu" and np.array_equal(o.numpy(), ref)


@pytest.mark.parametrize("mode", ["random", "clustered"])
def test_detect_batch64_headline_config(layers, mode):
    """BASELINE config 2: B=64 @640x640, conf 0.05, nms 0.3, top_k 750, nms_top_k 5000."""
    pri = synth.priors_numpy(640, 640)
    loc, conf = synth.detect_inputs(64, pri, 20262, 0.05, mode)
    ours = run_detect(layers, loc, conf, pri)
    assert_same(ours, oracle_detect(loc, conf, pri))
    again = run_detect(layers, loc, conf, pri)
    assert_same(ours, again)                                  # deterministic


def test_detect_1024_radix_select_path(layers):
    """N = 87,360 (config 5): ~19k candidates/image > the 8192-key sort capacity -> radix select."""
    pri = synth.priors_numpy(1024, 1024)
    loc, conf = synth.detect_inputs(3, pri, 77, 0.05, "random")
    assert (conf[..., 1] > 0.05).sum(1).min() > 8192
    assert_same(run_detect(layers, loc, conf, pri), oracle_detect(loc, conf, pri))


def test_detect_all_priors_are_candidates(layers):
    pri = synth.priors_numpy(640, 640)
    loc, conf = synth.detect_inputs(2, pri, 78, 0.0, "random")
    args = (2, 0, 750, 0.0, 0.3)
    assert_same(run_detect(layers, loc, conf, pri, args), oracle_detect(loc, conf, pri, args))


@pytest.mark.parametrize("args,nms_top_k", [((2, 0, 750, 0.3, 0.5), None), ((2, 0, 750, 0.2, 0.35), None),
                                            ((2, 0, 10, 0.05, 0.3), None), ((2, 0, 200, 0.05, 0.3), 300),
                                            ((2, 0, 750, 0.05, 0.3), 6000)])
def test_detect_production_thresholds(layers, args, nms_top_k):
    pri = synth.priors_numpy(640, 480)
    loc, conf = synth.detect_inputs(3, pri, 79, args[3], "random")
    assert_same(run_detect(layers, loc, conf, pri, args, nms_top_k=nms_top_k),
                oracle_detect(loc, conf, pri, args, nms_top_k=nms_top_k))


def test_detect_nms_top_k_beyond_limit_is_refused(layers):
    pri = synth.priors_numpy(640, 640)
    loc, conf = synth.detect_inputs(1, pri, 1, 0.05)
    with pytest.raises(NotImplementedError, match="nms_top_k"):
        run_detect(layers, loc, conf, pri, nms_top_k=8192)


def test_detect_three_classes(layers):
    pri = synth.priors_numpy(320, 320)
    rng = np.random.Generator(np.random.PCG64(5))
    N = pri.shape[0]
    loc = (rng.standard_normal((2, N, 4)) * 0.5).astype(np.float32)
    logits = rng.standard_normal((2, N, 3)) * 2 + np.array([3.0, 0, 0])
    conf = np.exp(logits); conf = (conf / conf.sum(-1, keepdims=True)).astype(np.float32)
    args = (3, 0, 100, 0.1, 0.4)
    assert_same(run_detect(layers, loc, conf, pri, args), oracle_detect(loc, conf, pri, args))


def test_detect_stage_entry_points_and_candidate_counts(layers):
    """stage 1 + stage 2 == fdt_detect; stage 1's candidate counts are exactly (score > conf_thresh).sum()."""
    from fdt_b200 import _lib
    pri = synth.priors_numpy(640, 480)
    loc, conf = synth.detect_inputs(5, pri, 321, 0.05)
    B, N, C = 5, pri.shape[0], 2
    dev = torch.device("cuda", torch.cuda.current_device())
    l, c, p = cu(loc), cu(conf), cu(pri)
    L = _lib.lib()
    ws = torch.empty(L.fdt_detect_workspace_bytes(B, N, C), dtype=torch.uint8, device=dev)
    st = _lib.stream_ptr()
    _lib.check(L.fdt_detect_threshold_compact(c.data_ptr(), B, N, C, 0.05, ws.data_ptr(), ws.numel(), st))
    cnt = torch.empty(B, dtype=torch.int32, device=dev)
    _lib.check(L.fdt_detect_candidate_counts(ws.data_ptr(), ws.numel(), B, N, C, cnt.data_ptr(), st))
    assert np.array_equal(npy(cnt), (conf[..., 1] > np.float32(0.05)).sum(1).astype(np.int32))
    out = torch.empty((B, C, 750, 5), dtype=torch.float32, device=dev)
    _lib.check(L.fdt_detect_sort_nms(l.data_ptr(), p.data_ptr(), B, N, C, 750, 5000, 0.3, 0.1, 0.2, out.data_ptr(), None, None,
                                     ws.data_ptr(), ws.numel(), st))
    assert np.array_equal(npy(out), oracle_detect(loc, conf, pri)[0])


def test_detect_rejects_nonpositive_nms_thresh(layers):
    with pytest.raises(ValueError):
        layers.Detect(2, 0, 750, 0.05, 0)


def test_detect_flat_reference_shapes(layers, golden):
    """Callers may pass loc as [B, N*4] and conf as [B*N, C] (detection.py docstring)."""
    g = golden("detect")
    loc, conf, pri = g["small_loc"], g["small_conf"], g["small_priors"]
    det = layers.Detect(2, 0, 750, 0.3, 0.5)
    out = det(cu(loc.reshape(4, -1)), cu(conf.reshape(-1, 2)), cu(pri))
    np.testing.assert_allclose(npy(out), g["small_out"], rtol=RTOL, atol=1e-7)


# ------------------------------------------------------------------------------- adversarial inputs for the
# bucket sort (score distributions) and the spatial grid (box geometry); the oracle defines the answer
def _rand_boxes(rng, n, lo=0.0, hi=1.0, smin=0.02, smax=0.15):
    ctr = rng.uniform(lo, hi, (n, 2)); wh = rng.uniform(smin, smax, (n, 2)) * (hi - lo)
    return np.concatenate([ctr - wh / 2, ctr + wh / 2], 1).astype(np.float32)


def _check_nms(layers, boxes, scores, thr, topk):
    keep, count = layers.box_utils.nms(cu(boxes), cu(scores), thr, topk)
    rk, rc = orc.nms(boxes, scores, thr, topk)
    assert count == rc, (count, rc)
    assert np.array_equal(npy(keep), rk)


@pytest.mark.parametrize("case", ["near_equal", "all_equal", "wide_range", "negative", "two_values", "one_bucket_overflow"])
def test_nms_score_distributions(layers, case):
    rng = np.random.Generator(np.random.PCG64(sum(map(ord, case))))
    n = 6000
    boxes = _rand_boxes(rng, n)
    if case == "near_equal":        # distinct scores a few ulps apart: one bucket holds everything -> radix-select fallback
        scores = (np.float32(0.5) + np.arange(n, dtype=np.float32) * np.float32(6e-8)).astype(np.float32)
        scores = rng.permutation(scores)
    elif case == "all_equal":       # ties resolved by index (higher first)
        scores = np.full(n, 0.7, np.float32)
    elif case == "wide_range":
        scores = np.exp(rng.uniform(-60, 60, n)).astype(np.float32)
    elif case == "negative":
        scores = rng.standard_normal(n).astype(np.float32) * 100
    elif case == "two_values":
        scores = rng.choice(np.array([0.25, 0.75], np.float32), n)
    else:                           # 5900 keys share one bucket, 100 spread far above
        scores = np.concatenate([np.float32(0.1) + np.arange(5900, dtype=np.float32) * np.float32(1e-8),
                                 rng.uniform(1, 1000, 100).astype(np.float32)]).astype(np.float32)
        scores = rng.permutation(scores)
    _check_nms(layers, boxes, scores, 0.4, 5000)


@pytest.mark.parametrize("case", ["pixels", "identical", "huge_and_tiny", "zero_area_far_apart", "inverted", "nan_inf",
                                  "thin", "tiny_thr", "thr_above_one", "dense_cells", "negative_coords", "single_point"])
def test_nms_box_geometry(layers, case):
    rng = np.random.Generator(np.random.PCG64(sum(map(ord, case))))
    n, thr, topk = 3000, 0.3, 5000
    boxes = _rand_boxes(rng, n)
    if case == "pixels":
        boxes = _rand_boxes(rng, n, 0, 1920, 0.01, 0.1)
    elif case == "identical":
        boxes = np.tile(np.array([[0.2, 0.2, 0.5, 0.6]], np.float32), (n, 1)); boxes[::7] += np.float32(0.31)
    elif case == "huge_and_tiny":
        boxes[:300] = _rand_boxes(rng, 300, 0, 1, 0.6, 1.5); boxes[300:900] = _rand_boxes(rng, 600, 0, 1, 1e-4, 1e-3)
    elif case == "zero_area_far_apart":       # 0/0 IoU = NaN suppresses at any distance (box_utils.py:337-339)
        pts = rng.uniform(0, 1, (200, 2)).astype(np.float32); boxes[:200] = np.concatenate([pts, pts], 1)
        boxes[200:260, 2] = boxes[200:260, 0]                                  # zero width, positive height
    elif case == "inverted":
        boxes[:100, [0, 2]] = boxes[:100, [2, 0]]
    elif case == "nan_inf":
        boxes[5, 1] = np.nan; boxes[77, 2] = np.inf; boxes[300, 0] = -np.inf; boxes[1500] = np.nan
    elif case == "thin":
        boxes[:1500, 3] = boxes[:1500, 1] + np.float32(1e-4); boxes[1500:, 2] = boxes[1500:, 0] + np.float32(0.9)
    elif case == "tiny_thr":
        thr = 0.004
    elif case == "thr_above_one":
        thr = 1.5
    elif case == "dense_cells":               # thousands of small boxes inside one grid cell
        boxes = _rand_boxes(rng, n, 0.50, 0.52, 0.01, 0.2); boxes[0] = [0, 0, 1, 1]
    elif case == "negative_coords":
        boxes = _rand_boxes(rng, n, -500, -100, 0.01, 0.1)
    elif case == "single_point":
        boxes = np.tile(np.array([[3, 3, 3, 3]], np.float32), (n, 1))
    scores = rng.permutation(np.unique(rng.uniform(0, 1, 3 * n).astype(np.float32)))[:n].copy()
    _check_nms(layers, boxes, scores, thr, topk)


def test_detect_mixed_pyramid_levels_and_scores(layers):
    """every pyramid level contributes candidates (boxes from 16 px to 512 px interact across grid levels)."""
    pri = synth.priors_numpy(640, 640)
    rng = np.random.Generator(np.random.PCG64(4242))
    N = pri.shape[0]
    loc = (rng.standard_normal((4, N, 4)) * 1.5).astype(np.float32)          # strong jitter: sizes vary by e^{+-0.3}
    s1 = rng.uniform(0, 1, (4, N)).astype(np.float32)
    s1[:, :25600] *= (rng.uniform(0, 1, (4, 25600)) < 0.1)                   # thin out level 0 so coarse levels matter
    conf = np.stack([1 - s1, s1], -1).astype(np.float32)
    for args in ((2, 0, 750, 0.05, 0.3), (2, 0, 750, 0.5, 0.7), (2, 0, 300, 0.05, 0.05)):
        assert_same(run_detect(layers, loc, conf, pri, args), oracle_detect(loc, conf, pri, args))


def test_detections_to_rows_consumer(layers, golden):
    """Detect's consumer loop (My_test.py:43-57): leading rows with score >= thr, fp32 scaling to pixels, per class."""
    from fdt_b200.utils.readout import detections_to_rows
    g = golden("detect")
    out = g["small_out"].copy()                         # [4, 2, 750, 5], images 1 and 2 empty
    out[0, 1, 5, 0] = 0.1                               # a row below the threshold stops the loop: later rows are not read
    for thr, w, h in ((0.3, 640.0, 480.0), (0.0, 640.0, 640.0), (0.9, 100.0, 50.0)):
        got = detections_to_rows(cu(out), thr, w, h, dummy_if_empty=True)
        scale = np.array([w, h, w, h], np.float32)
        for b in range(out.shape[0]):
            rows = []
            for i in range(out.shape[1]):
                j = 0
                while j < out.shape[2] and out[b, i, j, 0] >= np.float32(thr):      # verbatim loop of the reference
                    rows.append(np.concatenate([out[b, i, j, 1:] * scale, out[b, i, j, :1]]))
                    j += 1
            ref = np.array(rows, np.float32).reshape(-1, 5) if rows else np.array([[0, 0, 0, 0, 0.4]])
            assert got[b].dtype == ref.dtype and np.array_equal(got[b], ref)


def test_detect_cuda_graph_capture_and_side_stream(layers):
    """fdt_detect is asynchronous on the caller's stream, allocation-free and CUDA-graph capturable."""
    pri = synth.priors_numpy(640, 640)
    loc, conf = synth.detect_inputs(8, pri, 2024, 0.05)
    l, c, p = cu(loc), cu(conf), cu(pri)
    det = layers.Detect(2, 0, 750, 0.05, 0.3)
    ref = oracle_detect(loc, conf, pri)[0]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            out = det(l, c, p)                      # warm-up on the side stream (workspace allocation happens here)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            out = det(l, c, p)
    torch.cuda.synchronize()
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert np.array_equal(npy(out), ref)
    c2 = c.clone(); c2[:, :, 1] = 0.0                # new inputs in place: the replay must see them
    c.copy_(c2)
    g.replay()
    torch.cuda.synchronize()
    assert not npy(out).any()


@pytest.mark.parametrize("tag", ["1024", "480", "480c"])
def test_detect_golden_other_shapes(layers, golden, tag):
    """Reference fixtures at 1024x1024 (BASELINE config 5 shape) and at the tracker's 640x480 prior set (production thresholds)."""
    g = golden("detect_shapes")
    w, h, seed, B = (int(v) for v in g[tag + "_cfg"])
    a = g[tag + "_args"]
    args = (int(a[0]), int(a[1]), int(a[2]), float(a[3]), float(a[4]))
    pri = synth.priors_numpy(w, h)
    loc, conf = synth.detect_inputs(B, pri, seed, args[3], str(g[tag + "_mode"]))
    assert synth.digest(loc, conf) == str(g[tag + "_in_sha"])
    det = layers.Detect(*args)
    out, counts, kept = det(cu(loc), cu(conf), cu(pri), return_aux=True)
    out, counts, kept = npy(out), npy(counts), npy(kept)
    assert np.array_equal(counts, g[tag + "_counts"]) and np.array_equal(kept, g[tag + "_kept"])      # the reference's kept priors
    assert np.array_equal(out[:, 1, :, 0], g[tag + "_out"][..., 0])
    np.testing.assert_allclose(out[:, 1, :, 1:], g[tag + "_out"][..., 1:], rtol=RTOL, atol=1e-7)
    ref = orc.Detect(*args)(loc, conf, pri, return_aux=True)
    assert_same((out, counts, kept), ref)

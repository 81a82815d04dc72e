"""Sibling decode + NMS implementations: FaceBoxes DataEncoder (FACEBOX/encoderl.py) and MTCNN nms
(MTCNN/mtcnn/core/utils.py, core/nms.py).

CPU part: the oracle against tests/golden/siblings.npz (reference code run by oracle/make_golden.py gen_siblings).
GPU part: the CUDA path through the C ABI against the oracle (bit-exact indices) and the same fixture."""
import numpy as np
import pytest
import torch

from fdt_b200 import synth
from oracle import oracle as orc

S, M, P1, LE = orc.NMS_SUMFIRST, orc.NMS_MINIMUM, orc.NMS_PLUS1, orc.NMS_LE
FIXTURE_CASES = [            # fixture key, threshold, flags, reference function
    ("mt_union_06", 0.6, S), ("mt_min_04", 0.4, M), ("mt_min_05", 0.5, M),
    ("fb_np_union_05", 0.5, S), ("fb_np_min_03", 0.3, M),
    ("mt_plus1_union_05", 0.5, S | P1 | LE), ("mt_plus1_min_07", 0.7, M | P1 | LE),
    ("fb_torch_05", 0.5, S | LE), ("fb_torch_1", 1.0, S | LE),
]


def fb_inputs(g):
    rng = np.random.Generator(np.random.PCG64(51))
    N = 21824
    loc = (rng.standard_normal((N, 4)) * 0.5).astype(np.float32)
    s1 = 1.0 / (1.0 + np.exp(-(rng.standard_normal(N) * 2.0 - 4.0)))
    s1 = synth._uniquify_candidates(s1.astype(np.float32)[None], 0.35)[0]
    conf = np.stack([1 - s1, s1], 1).astype(np.float32)
    assert synth.digest(loc, conf) == str(g["fb_in_sha"])
    return loc, conf


def random_dets(n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    c = rng.uniform(20, 460, (max(n // 20, 1), 2))
    ctr = c[rng.integers(0, c.shape[0], n)] + rng.normal(0, 8, (n, 2))
    wh = rng.uniform(6, 120, (n, 2))
    sc = rng.permutation(n).astype(np.float64) / max(n, 1)
    return np.concatenate([ctr - wh / 2, ctr + wh / 2, sc[:, None]], 1).astype(np.float32)


# ------------------------------------------------------------------------------------------- oracle vs reference
def test_oracle_facebox_default_boxes(golden):
    g = golden("siblings")
    db = orc.facebox_default_boxes()
    assert db.shape == (21824, 4) and synth.digest(db) == str(g["fb_default_sha"])
    assert np.array_equal(db[:40], g["fb_default_head"]) and np.array_equal(db[-10:], g["fb_default_tail"])


@pytest.mark.parametrize("key,thr,flags", FIXTURE_CASES)
def test_oracle_nms_variants_match_reference(golden, key, thr, flags):
    g = golden("siblings")
    d = g["dets"]
    assert np.array_equal(orc.nms_variant(d[:, :4], d[:, 4], thr, flags), g[key])


def test_oracle_nms_variants_nan(golden):
    g = golden("siblings")
    d = g["dets_nan"]
    assert np.array_equal(orc.nms_variant(d[:, :4], d[:, 4], 0.5, S), g["nan_mt_union_05"])
    assert np.array_equal(orc.nms_variant(d[:, :4], d[:, 4], 0.5, M), g["nan_mt_min_05"])
    assert np.array_equal(orc.nms_variant(d[:, :4], d[:, 4], 0.5, S | P1 | LE), g["nan_plus1_union_05"])


def test_oracle_facebox_decode_np(golden):
    g = golden("siblings")
    loc, conf = fb_inputs(g)
    db = orc.facebox_default_boxes()
    boxes, scores = orc.facebox_decode_np(loc, conf, db)
    assert boxes.shape[0] == g["fb_kept"].shape[0]
    np.testing.assert_allclose(boxes, g["fb_boxes"], rtol=1e-5, atol=1e-7)        # np.exp vs fp64-rounded exp
    assert np.array_equal(scores, g["fb_scores"])


# ------------------------------------------------------------------------------------------- GPU vs oracle / reference
@pytest.mark.gpu
@pytest.mark.parametrize("key,thr,flags", FIXTURE_CASES)
def test_gpu_nms_variants_fixture(golden, key, thr, flags):
    from fdt_b200.siblings._nms import nms_variant
    g = golden("siblings")
    d = g["dets"]
    assert np.array_equal(nms_variant(d[:, :4], d[:, 4], thr, flags).cpu().numpy(), g[key])


@pytest.mark.gpu
def test_gpu_reference_signatures(golden):
    from fdt_b200.siblings import faceboxes, mtcnn
    g = golden("siblings")
    d = g["dets"]
    assert mtcnn.nms(d, 0.6, "Union") == g["mt_union_06"].tolist()
    assert mtcnn.nms(d, 0.4, "Minimum") == g["mt_min_04"].tolist()
    assert mtcnn.torch_nms(d, 0.7, "Minimum") == g["mt_plus1_min_07"].tolist()
    enc = faceboxes.DataEncoder()
    assert synth.digest(enc.default_boxes_np) == str(g["fb_default_sha"])
    assert enc.nms_np(d[:, :4], d[:, 4], 0.3, "Minimum") == g["fb_np_min_03"].tolist()
    k = enc.nms(torch.from_numpy(d[:, :4].copy()), torch.from_numpy(d[:, 4].copy()), 0.5)
    assert k.dtype == torch.int64 and k.tolist() == g["fb_torch_05"].tolist()
    dn = g["dets_nan"]
    assert mtcnn.nms(dn, 0.5, "Union") == g["nan_mt_union_05"].tolist()
    assert mtcnn.nms(dn, 0.5, "Minimum") == g["nan_mt_min_05"].tolist()
    assert mtcnn.torch_nms(dn, 0.5, "Union") == g["nan_plus1_union_05"].tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("flags", range(16))
@pytest.mark.parametrize("n,seed,thr", [(1, 1, 0.5), (2, 2, 0.5), (37, 3, 0.3), (900, 4, 0.5), (3000, 5, 0.7), (2500, 6, 0.05)])
def test_gpu_nms_variants_vs_oracle(flags, n, seed, thr):
    from fdt_b200.siblings._nms import nms_variant
    d = random_dets(n, 100 * seed + flags)
    if n >= 37:
        d[5, :4] = d[9, :4]                       # identical pair
        d[12, 2] = d[12, 0]                       # zero width
        d[20, 3] = d[20, 1] - 3.0                 # inverted
    ref = orc.nms_variant(d[:, :4], d[:, 4], thr, flags)
    got = nms_variant(d[:, :4], d[:, 4], thr, flags).cpu().numpy()
    assert np.array_equal(got, ref)


F64_CASES = [("f64_mt_union_06", 0.6, S), ("f64_mt_min_04", 0.4, M), ("f64_plus1_union_05", 0.5, S | P1 | LE), ("f64_fb_np_union_05", 0.5, S)]


@pytest.mark.parametrize("key,thr,flags", F64_CASES)
def test_oracle_nms_f64_matches_reference(golden, key, thr, flags):
    """float64 dets through the reference's own MTCNN / FaceBoxes NMS (the dtype its pipeline uses)."""
    g = golden("siblings")
    d = g["dets64"]
    assert d.dtype == np.float64 and np.array_equal(orc.nms_variant_f64(d[:, :4], d[:, 4], thr, flags), g[key])


@pytest.mark.gpu
@pytest.mark.parametrize("key,thr,flags", F64_CASES)
def test_gpu_nms_f64_fixture(golden, key, thr, flags):
    from fdt_b200.siblings import mtcnn
    from fdt_b200.siblings._nms import nms_variant
    g = golden("siblings")
    d = g["dets64"]
    assert np.array_equal(nms_variant(d[:, :4], d[:, 4], thr, flags, keep_dtype=True).cpu().numpy(), g[key])
    if key == "f64_mt_union_06":
        assert mtcnn.nms(d, 0.6, "Union") == g[key].tolist()                  # the drop-in keeps float64 dets in float64
    if key == "f64_plus1_union_05":
        assert mtcnn.torch_nms(d, 0.5, "Union") == g[key].tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [0, S, M, S | LE, S | P1 | LE, M | P1 | LE])
@pytest.mark.parametrize("n,seed,thr", [(1, 1, 0.5), (2, 2, 0.5), (65, 3, 0.3), (1000, 4, 0.5), (4100, 5, 0.6)])
def test_gpu_nms_f64_vs_oracle(flags, n, seed, thr):
    """float64 overlap rule on boxes that fp32 cannot represent, incl. degenerate boxes and an overlap that differs between the dtypes."""
    from fdt_b200.siblings._nms import nms_variant
    rng = np.random.Generator(np.random.PCG64(900 + seed))
    d = random_dets(n, 300 * seed + flags).astype(np.float64)
    d[:, :4] += rng.uniform(-1e-7, 1e-7, (n, 4))
    d[:, 4] += rng.uniform(0, 1e-10, n)
    if n >= 65:
        d[5, :4] = d[9, :4]; d[12, 2] = d[12, 0]; d[20, 3] = d[20, 1] - 3.0; d[30, 1] = np.nan
    ref = orc.nms_variant_f64(d[:, :4], d[:, 4], thr, flags)
    got = nms_variant(d[:, :4], d[:, 4], thr, flags, keep_dtype=True).cpu().numpy()
    assert np.array_equal(got, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("n,top_k,thr", [(9000, 0, 0.5), (20000, 12000, 0.3), (8500, 8400, 0.7)])
def test_gpu_nms_beyond_the_shared_memory_cap(n, top_k, thr):
    """layers/box_utils.nms has no candidate cap (box_utils.py:296-298): more than FDT_MAX_NMS_TOP_K candidates take the
    sort + pairwise-mask + reduce formulation; same keep list as the oracle, also for the sibling rules."""
    import torch
    from fdt_b200.layers import box_utils as bu
    from fdt_b200.siblings._nms import nms_variant
    pri = synth.priors_numpy(640, 640)
    rng = np.random.Generator(np.random.PCG64(n))
    sel = np.sort(rng.permutation(pri.shape[0])[:n])
    loc = (rng.standard_normal((n, 4)) * 0.5).astype(np.float32)
    boxes = orc.decode(loc, pri[sel], (0.1, 0.2))
    scores = rng.permutation(np.unique(rng.uniform(0.05, 1, 3 * n).astype(np.float32)))[:n].copy()
    keep, count = bu.nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), thr, top_k)
    rk, rc = orc.nms(boxes, scores, thr, top_k)
    assert int(count) == rc and np.array_equal(keep.cpu().numpy()[:rc], rk[:rc])
    assert not keep.cpu().numpy()[rc:].any()
    if n == 9000:
        assert np.array_equal(nms_variant(boxes, scores, thr, S | LE).cpu().numpy(), orc.nms_variant(boxes, scores, thr, S | LE))


@pytest.mark.gpu
@pytest.mark.parametrize("thr", [0.0, -1.0, 1.5, float("nan")])
def test_gpu_nms_variant_degenerate_thresholds(thr):
    from fdt_b200.siblings._nms import nms_variant
    d = random_dets(300, 77)
    for flags in (0, S, M, S | P1 | LE):
        assert np.array_equal(nms_variant(d[:, :4], d[:, 4], thr, flags).cpu().numpy(), orc.nms_variant(d[:, :4], d[:, 4], thr, flags))


@pytest.mark.gpu
def test_gpu_nms_variant_empty():
    from fdt_b200.siblings._nms import nms_variant
    assert nms_variant(np.zeros((0, 4), np.float32), np.zeros(0, np.float32), 0.5, S).numel() == 0


@pytest.mark.gpu
def test_gpu_facebox_decode_np(golden):
    from fdt_b200.siblings import faceboxes
    g = golden("siblings")
    loc, conf = fb_inputs(g)
    enc = faceboxes.DataEncoder()
    boxes, scores, idx = enc.decode_np(torch.from_numpy(loc), torch.from_numpy(conf), return_index=True)
    assert np.array_equal(idx, g["fb_kept"])                                       # the reference's kept default boxes
    assert np.array_equal(scores, g["fb_scores"])
    np.testing.assert_allclose(boxes, g["fb_boxes"], rtol=1e-5, atol=1e-7)
    rb, rs = orc.facebox_decode_np(loc, conf, orc.facebox_default_boxes())
    assert np.array_equal(boxes, rb) and np.array_equal(scores, rs)                # bit-exact against the oracle
    # lower threshold: many more candidates, heavy suppression between the density-tiled anchors
    for thr in (0.2, 0.6):
        b2, s2 = enc.decode_np(torch.from_numpy(loc), torch.from_numpy(conf), conf_thres=thr)
        rb, rs = orc.facebox_decode_np(loc, conf, orc.facebox_default_boxes(), conf_thres=thr)
        assert np.array_equal(b2, rb) and np.array_equal(s2, rs)


@pytest.mark.gpu
def test_gpu_facebox_decode_np_too_many_candidates(golden):
    from fdt_b200.siblings import faceboxes
    g = golden("siblings")
    loc, conf = fb_inputs(g)
    with pytest.raises(NotImplementedError):
        faceboxes.DataEncoder().decode_np(torch.from_numpy(loc), torch.from_numpy(conf), conf_thres=1e-4)


@pytest.mark.gpu
def test_gpu_facebox_encode(golden):
    from fdt_b200.siblings import faceboxes
    g = golden("siblings")
    enc = faceboxes.DataEncoder()
    gt = torch.from_numpy(g["enc_gt"])
    loc, conf = enc.encode(gt, torch.ones(gt.shape[0], dtype=torch.long))
    assert conf.dtype == torch.long and np.array_equal(conf.numpy(), g["enc_conf"])
    np.testing.assert_allclose(loc.numpy(), g["enc_loc"], rtol=1e-5, atol=1e-6)    # log: fp64-rounded vs torch.log
    assert np.array_equal(loc.numpy()[:, :2], g["enc_loc"][:, :2])                 # the centre offsets are plain fp32 arithmetic

"""Randomised (hypothesis) parity tests: many small ragged cases, CUDA path vs the CPU oracle, bit-exact.
Sizes are small so the whole file runs in seconds; the fixed-size tests cover the production shapes."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from fdt_b200 import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
SET = dict(deadline=None, max_examples=120, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture],
           derandomize=True)


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@settings(**SET)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(0, 700), thr=st.sampled_from([0.05, 0.3, 0.5, 0.9]),
       topk=st.sampled_from([1, 7, 200, 5000]), dup=st.booleans(), scale=st.sampled_from([1.0, 640.0, 1e-3]))
def test_nms_random(seed, n, thr, topk, dup, scale):
    import fdt_b200.layers as L
    rng = np.random.Generator(np.random.PCG64(seed))
    ctr = rng.uniform(0, 1, (n, 2)); wh = rng.uniform(0.0, 0.3, (n, 2))
    boxes = (np.concatenate([ctr - wh / 2, ctr + wh / 2], 1) * scale).astype(np.float32)
    scores = rng.uniform(0, 1, n).astype(np.float32)
    if dup and n > 4:
        scores[rng.integers(0, n, n // 3)] = scores[0]                 # equal scores: the tie rule decides
        boxes[rng.integers(0, n, n // 4)] = boxes[1]                   # identical boxes
    keep, count = L.box_utils.nms(cu(boxes), cu(scores), thr, topk)
    rk, rc = orc.nms(boxes, scores, thr, topk)
    assert count == rc and np.array_equal(keep.cpu().numpy(), rk)


@settings(**SET)
@given(seed=st.integers(0, 2**31 - 1), B=st.integers(1, 5), w=st.sampled_from([64, 96, 160]), h=st.sampled_from([64, 128]),
       conf_t=st.sampled_from([0.0, 0.05, 0.5, 0.99]), nms_t=st.sampled_from([0.1, 0.35, 0.7]), top_k=st.sampled_from([1, 20, 750]),
       nms_top_k=st.sampled_from([3, 100, 5000]))
def test_detect_random(seed, B, w, h, conf_t, nms_t, top_k, nms_top_k):
    import fdt_b200.layers as L
    pri = synth.priors_numpy(w, h)
    loc, conf = synth.detect_inputs(B, pri, seed, conf_t)
    ours = L.Detect(2, 0, top_k, conf_t, nms_t); ours.nms_top_k = nms_top_k
    ref = orc.Detect(2, 0, top_k, conf_t, nms_t); ref.nms_top_k = nms_top_k
    o, c, k = ours(cu(loc), cu(conf), cu(pri), return_aux=True)
    ro, rc, rk = ref(loc, conf, pri, return_aux=True)
    assert np.array_equal(c.cpu().numpy(), rc) and np.array_equal(k.cpu().numpy(), rk) and np.array_equal(o.cpu().numpy(), ro)


@settings(**SET)
@given(seed=st.integers(0, 2**31 - 1), G=st.integers(1, 40), bip=st.booleans(), thr=st.sampled_from([0.2, 0.35, 0.5]),
       w=st.sampled_from([64, 160]))
def test_match_random(seed, G, bip, thr, w):
    from tests.test_multibox_gpu import gpu_match
    rng = np.random.Generator(np.random.PCG64(seed))
    pri = synth.priors_numpy(w, w)
    gt = synth.gt_boxes(G, rng)
    if G > 2:
        gt[1] = gt[0]                                                   # duplicated GT: first wins the argmax, last the bipartite override
    lt, ct, bti, bto = gpu_match(bip, thr, gt, pri)
    o_lt, o_ct, o_bti, o_bto = orc.match(bip, thr, gt[:, :4], pri, (0.1, 0.2), gt[:, 4])
    assert np.array_equal(ct, o_ct) and np.array_equal(bti, o_bti) and np.array_equal(bto, o_bto)
    np.testing.assert_allclose(lt, o_lt, rtol=1e-6, atol=1e-7)


@settings(**SET)
@given(seed=st.integers(0, 2**31 - 1), B=st.integers(1, 4), N=st.integers(2, 3000), ratio=st.integers(0, 5), quant=st.booleans())
def test_mining_random(seed, B, N, ratio, quant):
    from tests.test_multibox_gpu import gpu_mine
    rng = np.random.Generator(np.random.PCG64(seed))
    lc = rng.uniform(0, 4, (B, N)).astype(np.float32)
    if quant:
        lc = (np.round(lc * 3) / 3).astype(np.float32)                  # massive ties
    pos = rng.uniform(0, 1, (B, N)) < rng.uniform(0, 0.3)
    lc[pos] = 0
    assert np.array_equal(gpu_mine(lc, pos, ratio), orc.hard_negative_mine(lc, pos, ratio))


@settings(**SET)
@given(seed=st.integers(0, 2**31 - 1), F=st.integers(1, 60), dmax=st.integers(1, 40), sigma=st.sampled_from([0.5, 3.0, 12.0]),
       s_iou=st.sampled_from([0.1, 0.4, 0.8]), s_h=st.sampled_from([0.0, 0.6]), t_min=st.sampled_from([0, 2, 5]),
       empty_every=st.sampled_from([0, 3, 11]))
def test_tracker_random(seed, F, dmax, sigma, s_iou, s_h, t_min, empty_every):
    from fdt_b200 import tracker as T
    frames = synth.tracker_frames(F=F, seed=seed, d_lo=1, d_hi=dmax, n_objects=dmax, empty_every=empty_every, sigma=sigma)
    a, b = T.iou_track(frames, s_iou, s_h, t_min), orc.iou_track(frames, s_iou, s_h, t_min)
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x["start_frame"] == y["start_frame"] and x["max_score"] == y["max_score"] and x["bboxes"] == y["bboxes"]


@pytest.mark.parametrize("which", [0, 1], ids=["exp", "log"])
def test_lean_exp_log_equal_the_math_library_on_every_float(which):
    """fdt_expf_cr / fdt_logf_cr (csrc/fdt_common.cuh): the lean fp64 evaluation of the common ranges must give the float that
    (float)exp((double)x) / (float)log((double)x) gives for ALL 2^32 bit patterns (decode, encode, log_sum_exp and the mining
    order of MultiBoxLoss depend on it bit for bit)."""
    from fdt_b200 import _lib
    dev = torch.device("cuda", torch.cuda.current_device())
    res = torch.zeros(2, dtype=torch.int64, device=dev)
    total = 0
    for first in range(0, 1 << 32, 1 << 30):
        _lib.check(_lib.lib().fdt_selftest_cr_math(which, first, 1 << 30, res.data_ptr(), _lib.stream_ptr()))
        bad, lowest = (int(v) for v in res.cpu().numpy().view(np.uint64))
        assert bad == 0, f"{bad} mismatching inputs from pattern {lowest - 1:#010x} on"
        total += 1 << 30
    assert total == 1 << 32

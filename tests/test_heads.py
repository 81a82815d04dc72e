"""Head post-processing that feeds Detect (pyramid.py:291-309, 331-332).

CPU part: the oracle against tests/golden/heads.npz (reference torch code, oracle/make_golden.py gen_heads).
GPU part: fdt_heads_to_loc_conf / fdt_detect_heads through the C ABI against the oracle, bit-exact."""
import numpy as np
import pytest
import torch

from fdt_b200 import synth
from oracle import oracle as orc

CASES = {"a": (synth.STRIDES6, synth.BOXES6), "b": (synth.STRIDES6[:5], synth.BOXES6[:5])}
SOFTMAX_RTOL = 1e-6       # exp: fp64 rounded once here, SLEEF / expf in torch (each within 1 ulp)


def case_inputs(g, tag):
    B, w, h, seed, nl = (int(v) for v in g[tag + "_cfg"])
    strides, boxes = CASES[tag]
    loc_maps, conf_maps, neg_max = synth.head_maps(B, w, h, seed, strides)
    if tag == "a":
        conf_maps[1][0, 2, 3, 4] = np.nan
        conf_maps[0][1, 0, 5, 6] = np.inf
    assert synth.digest(*loc_maps, *conf_maps) == str(g[tag + "_in_sha"])
    layer = orc.PriorBoxLayer(w, h, stride=strides, box=boxes)
    pri = np.concatenate([layer(i, fw, fh) for i, (fw, fh) in enumerate(synth.feature_maps(w, h, strides))], 0)
    return loc_maps, conf_maps, neg_max, pri


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_heads_match_reference(golden, tag):
    g = golden("heads")
    loc_maps, conf_maps, neg_max, pri = case_inputs(g, tag)
    loc, conf = orc.heads_to_loc_conf(loc_maps, conf_maps, neg_max, softmax=True)
    _, raw = orc.heads_to_loc_conf(loc_maps, conf_maps, neg_max, softmax=False)
    assert synth.digest(loc) == str(g[tag + "_loc_sha"])              # permute / cat: bit-exact
    assert synth.digest(raw) == str(g[tag + "_raw_sha"])              # max-in-out incl. NaN propagation: bit-exact
    ref = g[tag + "_conf"]
    assert np.array_equal(np.isnan(conf), np.isnan(ref))
    np.testing.assert_allclose(conf, ref, rtol=SOFTMAX_RTOL, atol=1e-30, equal_nan=True)
    # Detect on the oracle's conf keeps the same priors as the reference's Detect on the reference's conf
    out = orc.Detect(2, 0, 200, 0.05, 0.3)(loc, conf, pri)
    np.testing.assert_allclose(out, g[tag + "_out"], rtol=1e-5, atol=1e-7)
    assert np.array_equal(out[..., 0] > 0, g[tag + "_out"][..., 0] > 0)


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_gpu_heads_to_loc_conf_bit_exact(golden, tag):
    from fdt_b200.layers import heads_to_loc_conf
    g = golden("heads")
    loc_maps, conf_maps, neg_max, _ = case_inputs(g, tag)
    for softmax in (True, False):
        loc, conf = heads_to_loc_conf([cu(m) for m in loc_maps], [cu(m) for m in conf_maps], neg_max, softmax=softmax)
        rl, rc = orc.heads_to_loc_conf(loc_maps, conf_maps, neg_max, softmax=softmax)
        assert np.array_equal(loc.cpu().numpy(), rl)
        assert np.array_equal(conf.cpu().numpy(), rc, equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,thr", [("a", 0.05), ("b", 0.05), ("b", 0.3)])
def test_gpu_detect_heads_equals_detect_on_materialised(golden, tag, thr):
    from fdt_b200.layers import Detect, heads_to_loc_conf
    g = golden("heads")
    loc_maps, conf_maps, neg_max, pri = case_inputs(g, tag)
    det = Detect(2, 0, 200, thr, 0.3)
    lm, cm = [cu(m) for m in loc_maps], [cu(m) for m in conf_maps]
    out, counts, kept = det.detect_heads(lm, cm, cu(pri), neg_max, return_aux=True)
    rl, rc = orc.heads_to_loc_conf(loc_maps, conf_maps, neg_max, softmax=True)
    ref = orc.Detect(2, 0, 200, thr, 0.3)(rl, rc, pri)
    assert np.array_equal(out.cpu().numpy(), ref)                                      # fused path vs oracle
    out2, counts2, kept2 = det(*heads_to_loc_conf(lm, cm, neg_max), cu(pri), return_aux=True)
    assert torch.equal(out, out2) and torch.equal(counts, counts2) and torch.equal(kept, kept2)
    if thr == 0.05:
        assert np.array_equal(kept.cpu().numpy()[:, 1], g[tag + "_kept"][:, 1])       # the reference's kept priors


@pytest.mark.gpu
def test_gpu_detect_heads_headline_shape():
    """640x640, 6 levels, N = 34,125: fused path == Detect on the materialised tensors (bit-exact), B = 8."""
    from fdt_b200.layers import Detect, heads_to_loc_conf
    loc_maps, conf_maps, neg_max = synth.head_maps(8, 640, 640, 77)
    pri = synth.priors_numpy(640, 640)
    det = Detect(2, 0, 750, 0.05, 0.3)
    lm, cm = [cu(m) for m in loc_maps], [cu(m) for m in conf_maps]
    out = det.detect_heads(lm, cm, cu(pri))
    rl, rc = orc.heads_to_loc_conf(loc_maps, conf_maps, neg_max, softmax=True)
    ref = orc.Detect(2, 0, 750, 0.05, 0.3)(rl, rc, pri)
    assert np.array_equal(out.cpu().numpy(), ref)
    assert int((ref[:, 1, :, 0] > 0).sum()) > 8 * 100


@pytest.mark.gpu
def test_gpu_heads_argument_errors():
    from fdt_b200.layers import Detect, heads_to_loc_conf
    loc_maps, conf_maps, neg_max = synth.head_maps(1, 64, 64, 3)
    lm, cm = [cu(m) for m in loc_maps], [cu(m) for m in conf_maps]
    with pytest.raises(ValueError):
        heads_to_loc_conf(lm[:-1], cm)
    with pytest.raises(ValueError):
        heads_to_loc_conf(lm, [m[:, :2] for m in cm])
    with pytest.raises(ValueError):
        Detect(2, 0, 10, 0.05, 0.3).detect_heads(lm, cm, cu(synth.priors_numpy(64, 64))[:-1])
    with pytest.raises(NotImplementedError):
        Detect(3, 0, 10, 0.05, 0.3).detect_heads(lm, cm, cu(synth.priors_numpy(64, 64)))


@pytest.mark.gpu
@pytest.mark.parametrize("thr", [0.0, 1e-5, 0.5, 0.9999, 0.99995, 1.0, 1.5])
def test_gpu_detect_heads_threshold_range(golden, thr):
    """The logit-gap screening of k_heads_threshold_compact must stay a superset of the exact candidates at every threshold
    (the tiny / huge ones switch to the pass-all / fixed cut branches)."""
    from fdt_b200.layers import Detect
    g = golden("heads")
    loc_maps, conf_maps, neg_max, pri = case_inputs(g, "b")
    conf_maps = [m.copy() for m in conf_maps]
    conf_maps[0][0, 3, :8, :8] += 14.0                       # a block of very confident faces (scores that round to 1.0f)
    det = Detect(2, 0, 100, thr, 0.3)
    if thr <= 1e-5:
        det.nms_top_k = 8000                                 # nearly every prior is a candidate
    out, counts, kept = det.detect_heads([cu(m) for m in loc_maps], [cu(m) for m in conf_maps], cu(pri), neg_max, return_aux=True)
    rl, rc = orc.heads_to_loc_conf(loc_maps, conf_maps, neg_max, softmax=True)
    rdet = orc.Detect(2, 0, 100, thr, 0.3)
    rdet.nms_top_k = det.nms_top_k
    ref = rdet(rl, rc, pri)
    assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.gpu
def test_gpu_detect_heads_cuda_graph():
    from fdt_b200.layers import Detect
    loc_maps, conf_maps, neg_max = synth.head_maps(4, 320, 256, 9)
    pri = cu(synth.priors_numpy(320, 256))
    lm, cm = [cu(m) for m in loc_maps], [cu(m) for m in conf_maps]
    det = Detect(2, 0, 300, 0.05, 0.3)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        eager = det.detect_heads(lm, cm, pri)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            captured = det.detect_heads(lm, cm, pri)
    captured.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(captured, eager)

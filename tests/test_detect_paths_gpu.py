"""GPU parity tests for the launch paths of Detect that the default call does not exercise by itself: both k_sort_nms
instantiations (single CTA per list / 2-CTA cluster), batches with more lists than SMs, the overlap of consecutive calls through
the workspace ring (call sequencing in the workspace's control block), and the vectorised output stage at every alignment.
Everything is compared with the CPU oracle, bit for bit (layers/functions/detection.py:34-84, layers/box_utils.py:275-340)."""
import ctypes as C

import numpy as np
import pytest
import torch

from fdt_b200 import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def oracle_detect(loc, conf, pri, args=(2, 0, 750, 0.05, 0.3), nms_top_k=None):
    det = orc.Detect(*args); det.early_exit = True
    if nms_top_k:
        det.nms_top_k = nms_top_k
    return det(loc, conf, pri, return_aux=True)


@pytest.fixture()
def lib():
    from fdt_b200 import _lib
    yield _lib
    _lib.set_option("k3_cluster", -1)
    _lib.set_option("detect_depth", 4)
    _lib.set_option("detect_fused", -1)


def c_detect(_lib, l, c, p, ws, out, counts=None, kept=None, top_k=750, nms_top_k=5000, conf_t=0.05, nms_t=0.3):
    B, N, Cn = c.shape
    _lib.check(_lib.lib().fdt_detect(l.data_ptr(), c.data_ptr(), p.data_ptr(), B, N, Cn, top_k, nms_top_k, conf_t, nms_t, 0.1, 0.2,
                                     out.data_ptr(), _lib.ptr(counts), _lib.ptr(kept), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))


def new_ws(_lib, B, N, Cn, depth):
    return torch.empty(_lib.lib().fdt_detect_workspace_bytes_depth(B, N, Cn, depth), dtype=torch.uint8, device="cuda")


def status(_lib, ws):
    v = C.c_uint32(99)
    _lib.check(_lib.lib().fdt_detect_status(ws.data_ptr(), _lib.stream_ptr(), C.byref(v)))
    return v.value


# ------------------------------------------------------------------------------- both instantiations of k_sort_nms<MODE_DETECT>
VARIANTS = {"fused": (1, -1), "k2+single": (0, 0), "k2+cluster": (0, 1)}      # (detect_fused, k3_cluster)


def set_variant(lib, variant):
    fused, cluster = VARIANTS[variant]
    lib.set_option("detect_fused", fused)
    lib.set_option("k3_cluster", cluster)


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("mode", ["random", "clustered"])
def test_headline_batch_on_every_kernel_variant(lib, variant, mode):
    """B=64 @640x640 (BASELINE config 2) on the fused kernel (thresholds its own rows, the default), and on K2 followed by
    k_sort_nms<DETECT,1> (one CTA per list) and <DETECT,2> (2-CTA cluster)."""
    set_variant(lib, variant)
    cluster = variant
    pri = synth.priors_numpy(640, 640)
    loc, conf = synth.detect_inputs(64, pri, 20262, 0.05, mode)
    ref = oracle_detect(loc, conf, pri)
    l, c, p = cu(loc), cu(conf), cu(pri)
    for depth in (1, 3):
        ws = new_ws(lib, 64, pri.shape[0], 2, depth)
        out = torch.empty((64, 2, 750, 5), device="cuda")
        counts = torch.empty((64, 2), dtype=torch.int32, device="cuda")
        kept = torch.empty((64, 2, 750), dtype=torch.int64, device="cuda")
        c_detect(lib, l, c, p, ws, out, counts, kept)
        for got, want, name in zip((out, counts, kept), ref, ("out", "counts", "kept_prior")):
            assert np.array_equal(got.cpu().numpy(), want), f"{name} differs (cluster={cluster}, depth={depth})"


def test_more_lists_than_half_the_sms_uses_single_cta_path(lib):
    """B=80 at 160x160: lists * 2 > 148, so even the automatic policy with a one-slot workspace takes k_sort_nms<DETECT,1>."""
    import fdt_b200.layers as layers
    pri = synth.priors_numpy(160, 160)
    loc, conf = synth.detect_inputs(80, pri, 4242, 0.05)
    det = layers.Detect(2, 0, 750, 0.05, 0.3)
    got = det(cu(loc), cu(conf), cu(pri), return_aux=True)
    for g, w, name in zip(got, oracle_detect(loc, conf, pri), ("out", "counts", "kept_prior")):
        assert np.array_equal(g.cpu().numpy(), w), name
    ws = new_ws(lib, 80, pri.shape[0], 2, 1)
    out = torch.empty((80, 2, 750, 5), device="cuda")
    c_detect(lib, cu(loc), cu(conf), cu(pri), ws, out)
    assert np.array_equal(out.cpu().numpy(), oracle_detect(loc, conf, pri)[0])


def test_config5_shape_more_images_than_sms(lib):
    """BASELINE config 5's shape at a batch with more lists than SMs: B=150 @1024x1024 (N=87,360, ~19k candidates per image,
    unique scores): several waves of single-CTA lists, keys streamed from L2 (radix-select path)."""
    import fdt_b200.layers as layers
    pri = synth.priors_numpy(1024, 1024)
    loc, conf = synth.detect_inputs(150, pri, 99, 0.05)
    det = layers.Detect(2, 0, 750, 0.05, 0.3)
    got = det(cu(loc), cu(conf), cu(pri), return_aux=True)
    for g, w, name in zip(got, oracle_detect(loc, conf, pri), ("out", "counts", "kept_prior")):
        assert np.array_equal(g.cpu().numpy(), w), name


# ------------------------------------------------------------------------------- overlapped calls (workspace ring)
@pytest.mark.parametrize("depth", [1, 2, 3, 4])
@pytest.mark.parametrize("variant", list(VARIANTS))
def test_back_to_back_calls_overlap_and_stay_exact(lib, depth, variant):
    """16 calls issued back to back on one stream over different batches: with depth >= 2 they overlap on the device; every
    output must equal the oracle's, whether each call has its own output buffer or all share one (then the calls take turns)."""
    set_variant(lib, variant)
    pri = synth.priors_numpy(640, 640)
    N = pri.shape[0]
    n_sets, B = 4, 8
    data = [synth.detect_inputs(B, pri, 9000 + i, 0.05, "random" if i % 2 == 0 else "clustered") for i in range(n_sets)]
    refs = [oracle_detect(l, c, pri)[0] for l, c in data]
    dev = [(cu(l), cu(c)) for l, c in data]
    p = cu(pri)
    ws = new_ws(lib, B, N, 2, depth)
    calls = 16
    outs = [torch.empty((B, 2, 750, 5), device="cuda") for _ in range(calls)]
    torch.cuda._sleep(3_000_000)                     # the host enqueues all calls while the GPU spins: they reach the device back to back
    for i in range(calls):
        c_detect(lib, dev[i % n_sets][0], dev[i % n_sets][1], p, ws, outs[i])
    torch.cuda.synchronize()
    for i in range(calls):
        assert np.array_equal(outs[i].cpu().numpy(), refs[i % n_sets]), f"call {i} (depth {depth})"
    # one shared output buffer: a call must wait for the previous one; what remains is the LAST call's result
    shared = torch.empty((B, 2, 750, 5), device="cuda")
    copies = []
    torch.cuda._sleep(3_000_000)
    for i in range(7):
        c_detect(lib, dev[i % n_sets][0], dev[i % n_sets][1], p, ws, shared)
        if i in (2, 6):
            copies.append((i, shared.clone()))       # an ordinary kernel after the call: must see the call complete
    torch.cuda.synchronize()
    for i, t in copies:
        assert np.array_equal(t.cpu().numpy(), refs[i % n_sets]), f"shared output after call {i}"
    assert status(lib, ws) == 0


def test_inputs_produced_on_stream_between_overlapped_calls(lib):
    """Each call's conf / loc are written by ordinary kernels enqueued right before it (into one of four staging buffers) and each
    output is consumed by an ordinary kernel right after.  The producers must be complete before the call reads its inputs and
    the consumer must see the finished rows, while the calls themselves overlap where they can."""
    pri = synth.priors_numpy(640, 480)
    N = pri.shape[0]
    B = 6
    data = [synth.detect_inputs(B, pri, 7100 + i, 0.05) for i in range(3)]
    refs = [oracle_detect(l, c, pri)[0] for l, c in data]
    src = [(cu(l), cu(c)) for l, c in data]
    p = cu(pri)
    ws = new_ws(lib, B, N, 2, 3)
    stage_l = [torch.zeros_like(src[0][0]) for _ in range(4)]
    stage_c = [torch.zeros_like(src[0][1]) for _ in range(4)]
    outs = [torch.empty((B, 2, 750, 5), device="cuda") for _ in range(4)]
    got = []
    torch.cuda._sleep(3_000_000)
    for i in range(12):
        j = i % 4
        stage_l[j].copy_(src[i % 3][0]); stage_c[j].copy_(src[i % 3][1])      # producers (ordinary kernels)
        c_detect(lib, stage_l[j], stage_c[j], p, ws, outs[j])
        got.append(outs[j] * 1.0)                                             # consumer (ordinary kernel)
    torch.cuda.synchronize()
    for i, t in enumerate(got):
        assert np.array_equal(t.cpu().numpy(), refs[i % 3]), f"call {i}"


@pytest.mark.parametrize("variant", ["fused", "k2+single"])
def test_refilled_input_buffers_full_batch(lib, variant):
    """The same two conf / loc buffers are refilled with new contents before every call, at B = 64 (every SM runs CTAs of several
    calls that read the same addresses): no call may see what an earlier call read there (read-only / L1 cached loads), while the
    calls overlap wherever the sequencing allows it."""
    set_variant(lib, variant)
    pri = synth.priors_numpy(640, 640)
    N = pri.shape[0]
    B = 64
    data = [synth.detect_inputs(B, pri, 6100 + i, 0.05, "random" if i != 1 else "clustered") for i in range(3)]
    refs = [oracle_detect(l, c, pri)[0] for l, c in data]
    src = [(cu(l), cu(c)) for l, c in data]
    p = cu(pri)
    ws = new_ws(lib, B, N, 2, 4)
    stage_l = [torch.zeros_like(src[0][0]) for _ in range(2)]
    stage_c = [torch.zeros_like(src[0][1]) for _ in range(2)]
    outs = [torch.empty((B, 2, 750, 5), device="cuda") for _ in range(18)]
    torch.cuda._sleep(5_000_000)
    for i in range(18):
        j = i % 2
        stage_l[j].copy_(src[i % 3][0]); stage_c[j].copy_(src[i % 3][1])
        c_detect(lib, stage_l[j], stage_c[j], p, ws, outs[i])
    torch.cuda.synchronize()
    for i in range(18):
        assert np.array_equal(outs[i].cpu().numpy(), refs[i % 3]), f"call {i}"
    assert status(lib, ws) == 0


def test_stage1_alone_then_full_calls_and_geometry_change(lib):
    """A stage-1 call that never gets its stage 2 (tests and tools do that) must not wedge the sequencing; neither must a different
    batch geometry on the same workspace memory, nor memory that holds garbage."""
    pri = synth.priors_numpy(320, 320)
    N = pri.shape[0]
    loc, conf = synth.detect_inputs(5, pri, 55, 0.05)
    ref5 = oracle_detect(loc, conf, pri)[0]
    ref3 = oracle_detect(loc[:3], conf[:3], pri)[0]
    l, c, p = cu(loc), cu(conf), cu(pri)
    L = lib.lib()
    ws = new_ws(lib, 5, N, 2, 3)
    ws.random_(0, 256)                                                        # garbage, including the control block
    st = lib.stream_ptr()
    out = torch.empty((5, 2, 750, 5), device="cuda")
    for rep in range(3):
        lib.check(L.fdt_detect_threshold_compact(c.data_ptr(), 5, N, 2, 0.05, ws.data_ptr(), ws.numel(), st))      # dangling stage 1
        c_detect(lib, l, c, p, ws, out)
        assert np.array_equal(out.cpu().numpy(), ref5)
        out3 = torch.empty((3, 2, 750, 5), device="cuda")
        c_detect(lib, l[:3], c[:3], p, ws, out3)                              # other geometry, same memory, previous call in flight
        c_detect(lib, l, c, p, ws, out)
        torch.cuda.synchronize()
        assert np.array_equal(out3.cpu().numpy(), ref3) and np.array_equal(out.cpu().numpy(), ref5)
    assert status(lib, ws) == 0


def test_python_detect_calls_overlap_with_fresh_outputs(lib):
    """The drop-in Detect object owns a ring workspace per stream and shape: back-to-back calls, each with the fresh output tensor
    it allocates, stay exact; so do two streams using the same object."""
    import fdt_b200.layers as layers
    pri = synth.priors_numpy(640, 640)
    data = [synth.detect_inputs(4, pri, 8100 + i, 0.05) for i in range(3)]
    refs = [oracle_detect(l, c, pri)[0] for l, c in data]
    dev = [(cu(l), cu(c)) for l, c in data]
    p = cu(pri)
    det = layers.Detect(2, 0, 750, 0.05, 0.3)
    outs = [det(dev[i % 3][0], dev[i % 3][1], p) for i in range(9)]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        outs2 = [det(dev[i % 3][0], dev[i % 3][1], p) for i in range(6)]
    torch.cuda.synchronize()
    for i, o in enumerate(outs):
        assert np.array_equal(o.cpu().numpy(), refs[i % 3])
    for i, o in enumerate(outs2):
        assert np.array_equal(o.cpu().numpy(), refs[i % 3])
    assert det._workspaces.status() == 0


# ------------------------------------------------------------------------------- vectorised output stage
@pytest.mark.parametrize("top_k,nms_top_k", [(1, 5000), (3, 5000), (10, 5000), (750, 5000), (1601, 5000), (3300, 6000)])
@pytest.mark.parametrize("shift", [0, 1, 2, 3])
def test_output_rows_at_every_alignment(lib, top_k, nms_top_k, shift):
    """The [top_k, 5] planes are staged in shared memory and stored with 16-byte vector stores after a scalar lead-in: every
    phase of the destination address (shift floats past a 16-byte boundary), short planes, and planes longer than one staging
    chunk (1600 rows); the floats around the output must stay untouched."""
    pri = synth.priors_numpy(320, 320)
    N = pri.shape[0]
    B = 3
    loc, conf = synth.detect_inputs(B, pri, 31 + top_k, 0.01)
    args = (2, 0, top_k, 0.01, 0.45)
    ref = oracle_detect(loc, conf, pri, args, nms_top_k=nms_top_k)[0]
    n = B * 2 * top_k * 5
    buf = torch.full((n + 16,), -7.0, device="cuda")
    out = buf[4 + shift: 4 + shift + n]
    assert out.data_ptr() % 16 == 4 * shift
    ws = new_ws(lib, B, N, 2, 2)
    c_detect(lib, cu(loc), cu(conf), cu(pri), ws, out, top_k=top_k, nms_top_k=nms_top_k, conf_t=0.01, nms_t=0.45)
    got = buf.cpu().numpy()
    assert np.array_equal(got[4 + shift: 4 + shift + n].reshape(ref.shape), ref)
    assert np.all(got[:4 + shift] == -7.0) and np.all(got[4 + shift + n:] == -7.0)


# ------------------------------------------------------------------------------- fused kernel: score distributions and class counts
@pytest.mark.parametrize("variant", ["fused", "k2+single"])
@pytest.mark.parametrize("dist", ["ties", "tiny-range", "above-one", "negative-thresh", "nan-inf"])
def test_degenerate_score_distributions(lib, variant, dist):
    """The fused kernel buckets scores over the a-priori range (conf_thresh, 1]; everything that defeats the bucket map (thousands
    of equal scores, all scores within a few ulps, scores above 1, a negative threshold, NaN / inf scores) must fall back to the
    exact selection paths and still equal the oracle."""
    set_variant(lib, variant)
    pri = synth.priors_numpy(640, 480)
    N = pri.shape[0]
    B = 3
    rng = np.random.Generator(np.random.PCG64(77))
    loc, conf = synth.detect_inputs(B, pri, 500, 0.05)
    thr = 0.05
    s = conf[..., 1].copy()
    if dist == "ties":
        s[0, rng.permutation(N)[:9000]] = np.float32(0.7)          # 9,000 equal scores straddle rank nms_top_k
        s[1, :] = np.float32(0.25)                                 # every prior a candidate, all tied
    elif dist == "tiny-range":
        base = np.float32(0.3)
        s[:] = base + (rng.integers(0, 7, s.shape) * np.spacing(base)).astype(np.float32)
    elif dist == "above-one":
        s[0, ::3] = rng.uniform(1.0, 50.0, s[0, ::3].shape).astype(np.float32)
        s[1, ::5] = np.float32(1.0)
    elif dist == "negative-thresh":
        thr = -0.5
        s[:] = rng.uniform(-1.0, 1.0, s.shape).astype(np.float32)
    elif dist == "nan-inf":
        s[0, ::7] = np.nan; s[0, 1::7] = np.inf; s[1, ::9] = -np.inf; s[2, 5] = np.inf
    conf = np.stack([1 - s, s], -1).astype(np.float32)
    args = (2, 0, 750, thr, 0.3)
    ref = oracle_detect(loc, conf, pri, args)
    ws = new_ws(lib, B, N, 2, 2)
    out = torch.empty((B, 2, 750, 5), device="cuda")
    counts = torch.empty((B, 2), dtype=torch.int32, device="cuda")
    kept = torch.empty((B, 2, 750), dtype=torch.int64, device="cuda")
    c_detect(lib, cu(loc), cu(conf), cu(pri), ws, out, counts, kept, conf_t=thr)
    for got, want, name in zip((out, counts, kept), ref, ("out", "counts", "kept_prior")):
        assert np.array_equal(got.cpu().numpy(), want, equal_nan=True), f"{name} differs ({dist}, {variant})"


@pytest.mark.parametrize("variant", ["fused", "k2+single"])
def test_four_classes_and_quirk_lists(lib, variant):
    """C = 4: three lists per image read with stride C; lists with zero and with exactly one candidate (detection.py:66-72)."""
    set_variant(lib, variant)
    pri = synth.priors_numpy(320, 320)
    N = pri.shape[0]
    rng = np.random.Generator(np.random.PCG64(5))
    loc = (rng.standard_normal((3, N, 4)) * 0.5).astype(np.float32)
    logits = rng.standard_normal((3, N, 4)) * 2 + np.array([3.0, 0, 0, -1.0])
    conf = np.exp(logits); conf = (conf / conf.sum(-1, keepdims=True)).astype(np.float32)
    conf[1, :, 2] = 0.01; conf[1, 123, 2] = 0.8            # exactly one candidate
    conf[2, :, 3] = 0.0                                     # none
    args = (4, 0, 100, 0.1, 0.4)
    ref = oracle_detect(loc, conf, pri, args)
    ws = new_ws(lib, 3, N, 4, 3)
    out = torch.empty((3, 4, 100, 5), device="cuda")
    counts = torch.empty((3, 4), dtype=torch.int32, device="cuda")
    kept = torch.empty((3, 4, 100), dtype=torch.int64, device="cuda")
    c_detect(lib, cu(loc), cu(conf), cu(pri), ws, out, counts, kept, top_k=100, conf_t=0.1, nms_t=0.4)
    for got, want, name in zip((out, counts, kept), ref, ("out", "counts", "kept_prior")):
        assert np.array_equal(got.cpu().numpy(), want), name
    assert counts[1, 2].item() == 0 and counts[2, 3].item() == 0

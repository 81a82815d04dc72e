"""GPU tests of the host-buffer pipeline (fdt_detect_host / _submit / _wait, fdt_ctx_set_priors): chunked copy/compute overlap,
several calls in flight, resident priors, re-planning on a shape change -- every result bit-identical to the oracle's."""
import ctypes as C
from collections import deque

import numpy as np
import pytest
import torch

from fdt_b200 import _lib, synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
ARGS = (2, 0, 750, 0.05, 0.3)


@pytest.fixture(scope="module")
def layers():
    import fdt_b200.layers as L
    return L


@pytest.fixture(autouse=True)
def default_chunk():
    yield
    _lib.set_option("host_chunk", 16)


def oracle_detect(loc, conf, pri, args=ARGS):
    det = orc.Detect(*args); det.early_exit = True
    return det(loc, conf, pri, return_aux=True)


def pinned(*arrays):
    return [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in arrays]


def same(ours, ref):
    out, counts, kept = (t.numpy() for t in ours)
    assert np.array_equal(counts, ref[1])
    assert np.array_equal(kept, ref[2])
    assert np.array_equal(out, ref[0])


PRI = synth.priors_numpy(640, 640)


@pytest.mark.parametrize("chunk", [16, 0, 8, 5])
@pytest.mark.parametrize("pin", [True, False])
def test_chunked_call_with_remainder(layers, chunk, pin):
    """B = 40: chunks of 16 + 16 + 8 (a remainder with its own workspace), 8 x 5, 5 x 8, or one block -- same bits as the oracle."""
    _lib.set_option("host_chunk", chunk)
    loc, conf = synth.detect_inputs(40, PRI, 77, 0.05)
    det = layers.Detect(*ARGS)
    tensors = pinned(loc, conf, PRI) if pin else [torch.from_numpy(a) for a in (loc, conf, PRI)]
    ref = oracle_detect(loc, conf, PRI)
    for _ in range(2):                                         # second call: staging slots and workspaces are reused
        same(det(*tensors, return_aux=True), ref)


def test_calls_in_flight_complete_in_order(layers):
    """Eight batches, three in flight: every result equals the oracle's for ITS inputs (slots, events and the Detect ring are reused)."""
    det = layers.Detect(*ARGS)
    batches = [synth.detect_inputs(24, PRI, 500 + i, 0.05, "clustered" if i % 3 == 0 else "random") for i in range(8)]
    pri_t = pinned(PRI)[0]
    host = [pinned(l, c) for l, c in batches]
    pend, results = deque(), []
    for l, c in host:
        pend.append(det.submit(l, c, pri_t, return_aux=True))
        if len(pend) > 3:
            results.append(pend.popleft().result())
    while pend:
        results.append(pend.popleft().result())
    for (l, c), r in zip(batches, results):
        same(r, oracle_detect(l, c, PRI))
    # a result may be asked for twice, and out of order
    p1 = det.submit(*host[0], pri_t, return_aux=True); p2 = det.submit(*host[1], pri_t, return_aux=True)
    same(p2.result(), oracle_detect(*batches[1], PRI)); same(p1.result(), oracle_detect(*batches[0], PRI)); same(p1.result(), oracle_detect(*batches[0], PRI))


def test_prior_set_changes_between_calls_in_flight(layers):
    """The resident prior set follows the tensor the caller passes: A, B, B, A with all four calls in flight."""
    det = layers.Detect(*ARGS)
    loc, conf = synth.detect_inputs(8, PRI, 91, 0.05)
    pri_b = PRI.copy(); pri_b[:, 2:] *= np.float32(1.25)
    ta, tb = pinned(PRI)[0], pinned(pri_b)[0]
    l, c = pinned(loc, conf)
    pend = [det.submit(l, c, p, return_aux=True) for p in (ta, tb, tb, ta)]
    ra, rb = oracle_detect(loc, conf, PRI), oracle_detect(loc, conf, pri_b)
    assert not np.array_equal(ra[0], rb[0])
    for p, r in zip(pend, (ra, rb, rb, ra)):
        same(p.result(), r)
    # an in-place torch write bumps the version counter: the set is uploaded again
    ta[:, 2:] *= 1.25
    same(det(l, c, ta, return_aux=True), rb)


def test_shape_changes_replan_the_context(layers):
    det = layers.Detect(*ARGS)
    pri480 = synth.priors_numpy(640, 480)
    cases = [(4, PRI, 1), (20, PRI, 2), (3, pri480, 3), (4, PRI, 4), (33, pri480, 5)]
    for B, pri, seed in cases:
        loc, conf = synth.detect_inputs(B, pri, seed, 0.05)
        same(det(*pinned(loc, conf, pri), return_aux=True), oracle_detect(loc, conf, pri))
    det5 = layers.Detect(2, 0, 5, 0.05, 0.3)                  # top_k 5: 40-byte rows per list, chunk blocks stay 16-byte aligned only at even offsets
    _lib.set_option("host_chunk", 3)
    loc, conf = synth.detect_inputs(7, PRI, 6, 0.05)
    same(det5(*pinned(loc, conf, PRI), return_aux=True), oracle_detect(loc, conf, PRI, (2, 0, 5, 0.05, 0.3)))


def test_c_abi_tickets_and_missing_priors():
    L = _lib.lib()
    ctx = C.c_void_p()
    _lib.check(L.fdt_ctx_create(torch.cuda.current_device(), C.byref(ctx)))
    try:
        loc, conf = synth.detect_inputs(2, PRI, 3, 0.05)
        l, c, p = pinned(loc, conf, PRI)
        out = torch.empty((2, 2, 750, 5), dtype=torch.float32).pin_memory()
        t = C.c_uint64(0)
        tail = (2, PRI.shape[0], 2, 750, 5000, 0.05, 0.3, 0.1, 0.2, _lib.ptr(out), None, None)
        assert L.fdt_detect_host_wait(ctx, 1) != 0 and b"never issued" in L.fdt_last_error()
        assert L.fdt_detect_host_submit(ctx, _lib.ptr(l), _lib.ptr(c), None, *tail, C.byref(t)) != 0      # no resident set yet
        assert b"priors" in L.fdt_last_error()
        _lib.check(L.fdt_ctx_set_priors(ctx, _lib.ptr(p), PRI.shape[0]))
        _lib.check(L.fdt_detect_host_submit(ctx, _lib.ptr(l), _lib.ptr(c), None, *tail, C.byref(t)))
        assert t.value == 1
        _lib.check(L.fdt_detect_host_wait(ctx, 1))
        assert np.array_equal(out.numpy(), oracle_detect(loc, conf, PRI)[0])
        assert L.fdt_detect_host_wait(ctx, 2) != 0
        # an empty batch completes at once
        _lib.check(L.fdt_detect_host_submit(ctx, None, None, None, 0, PRI.shape[0], 2, 750, 5000, 0.05, 0.3, 0.1, 0.2, None, None, None, C.byref(t)))
        assert t.value == 2
        _lib.check(L.fdt_detect_host_wait(ctx, 2))
    finally:
        _lib.check(L.fdt_ctx_destroy(ctx))

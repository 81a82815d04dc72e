#!/usr/bin/env python
"""Secondary measurements for the other SURVEY section-8 rows (not the driver's headline; bench.py is):

    python bench_extra.py multibox   # BASELINE config 3: MultiBoxLoss B=32, N=34,125, G in [0,200]
    python bench_extra.py tracker    # BASELINE config 4: IoU tracker, 10k frames x 1-300 detections
    python bench_extra.py detect1024 # BASELINE config 5 shape on one GPU: Detect B=64 @1024^2 (N=87,360)
    python bench_extra.py priorbox

Each prints one JSON line with device time (CUDA events, L2 flushed between repetitions), the algorithmic bytes of
SURVEY 8(d) and the CPU oracle port timed on the host beside it."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import fdt_b200  # noqa: F401
from fdt_b200 import _lib, synth
from oracle import oracle as orc

PEAK = 6500.6
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, reps=20, warm=3):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        torch.cuda._sleep(250_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.mean(ts)), float(np.min(ts))


def cpu_time(fn, budget=8.0):
    fn()
    n, t0 = 0, time.perf_counter()
    while True:
        fn(); n += 1
        dt = time.perf_counter() - t0
        if dt > budget or n >= 20:
            return dt / n, n


def multibox():
    from fdt_b200.layers import MultiBoxLoss
    from fdt_b200.layers.modules.multibox_loss import pack_targets
    B = 32
    pri = synth.priors_numpy(640, 640); N = pri.shape[0]
    loc, conf, targets = synth.multibox_inputs(B, pri, 3030, 0, 200)
    dev = torch.device("cuda")
    l, c, p = (torch.from_numpy(a).to(dev) for a in (loc, conf, pri))
    tg = [torch.from_numpy(t).to(dev) for t in targets]
    out = {}
    for bip in (False, True):
        crit = MultiBoxLoss(2, 0.35, True, 0, True, 3, 0.35, False, bipartite=bip)
        gt, off, total = pack_targets(tg, dev)
        L = _lib.lib()
        losses = torch.empty(2, device=dev); norm = torch.empty(1, device=dev)
        loc_t = torch.empty((B, N, 4), device=dev); conf_t = torch.empty((B, N), dtype=torch.int64, device=dev)
        sel = torch.empty((B, N), dtype=torch.uint8, device=dev)
        ws = _lib.workspace(L.fdt_multibox_workspace_bytes(B, N, 2, total), dev, "bx")

        def fwd():
            _lib.check(L.fdt_multibox_loss_forward(l.data_ptr(), c.data_ptr(), p.data_ptr(), gt.data_ptr(), off.data_ptr(), total,
                                                   B, N, 2, 0.35, 3, int(bip), 0.1, 0.2, losses.data_ptr(), norm.data_ptr(),
                                                   loc_t.data_ptr(), conf_t.data_ptr(), sel.data_ptr(), None, ws.data_ptr(),
                                                   ws.numel(), _lib.stream_ptr()))
        ms, ms_min = timed(fwd)
        lr = l.clone().requires_grad_(True); cr = c.clone().requires_grad_(True)

        def fwdbwd():
            ll, lc = crit((lr, cr, p), tg)
            (ll + lc).backward()
        ms_fb, _ = timed(fwdbwd, reps=10)
        out["bipartite" if bip else "default"] = {"forward_ms": ms, "forward_ms_min": ms_min, "images_per_s": B / (ms * 1e-3),
                                                  "module_fwd_bwd_ms": ms_fb}
    G = sum(t.shape[0] for t in targets)
    alg = B * 48 * N + 20 * G
    cpu_s, n = cpu_time(lambda: orc.multibox_loss(loc, conf, pri, targets, 0.35, 3, False, want_aux=False))
    d = out["default"]
    print(json.dumps({"workload": "MultiBoxLoss match/encode + mining + loss, B=32, N=34,125, G~U{0..200} (config 3)", "results": out,
                      "algorithmic_bytes": alg, "achieved_gbs": alg / (d["forward_ms"] * 1e-3) / 1e9,
                      "roofline_frac_of_measured_hbm": alg / (d["forward_ms"] * 1e-3) / 1e9 / PEAK, "gt_total": G,
                      "iou_evaluations": int(sum(t.shape[0] for t in targets) * N),
                      "cpu_baseline": {"images_per_s": B / cpu_s, "cores": orc.max_threads(), "kind": "port", "sample": f"{n} batches"}}))


def tracker():
    from fdt_b200 import tracker as T
    frames = synth.tracker_frames(F=10000, seed=4040, d_lo=1, d_hi=300, n_objects=300, empty_every=1000)
    dets, off = T.pack_frames(frames)
    d = torch.from_numpy(dets).cuda(); o = torch.from_numpy(off).cuda()
    F = len(frames); total = int(off[-1]); max_d = int(np.diff(off).max())
    L = _lib.lib()
    dev = torch.device("cuda")
    n = torch.zeros(1, dtype=torch.int64, device=dev); t_off = torch.zeros(total + 2, dtype=torch.int64, device=dev)
    t_dets = torch.zeros(total, dtype=torch.int64, device=dev); t_start = torch.zeros(total + 1, dtype=torch.int64, device=dev)
    t_max = torch.zeros(total + 1, dtype=torch.float64, device=dev)
    ws = _lib.workspace(L.fdt_iou_track_workspace_bytes(F, total, max_d), dev, "tx")

    def run():
        _lib.check(L.fdt_iou_track(d.data_ptr(), o.data_ptr(), F, total, max_d, 0.4, 0.6, 5, n.data_ptr(), t_off.data_ptr(),
                                   t_dets.data_ptr(), t_start.data_ptr(), t_max.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    ms, ms_min = timed(run, reps=5, warm=1)
    t0 = time.perf_counter(); tr = T.iou_track(frames); api_s = time.perf_counter() - t0
    t0 = time.perf_counter(); ref = orc.iou_track(frames); cpu_s = time.perf_counter() - t0
    same = len(tr) == len(ref) and all(a["bboxes"] == b["bboxes"] and a["start_frame"] == b["start_frame"] for a, b in zip(tr, ref))
    print(json.dumps({"workload": "IoU tracker association, 10,000 frames x U{1..300} detections (config 4)", "frames": F,
                      "detections": total, "tracks": len(tr), "device_ms": ms, "device_ms_min": ms_min, "frames_per_s_device": F / (ms * 1e-3),
                      "frames_per_s_api_incl_python_dict_build": F / api_s, "identical_to_oracle": bool(same),
                      "algorithmic_bytes": 28 * total, "note": "latency-bound serial chain; roofline fraction not meaningful (SURVEY 8d)",
                      "cpu_baseline": {"frames_per_s": F / cpu_s, "cores": 1, "kind": "port", "sample": "whole video, C oracle, single thread"}}))


def detect1024():
    from fdt_b200.layers import Detect
    B = 64
    pri = synth.priors_numpy(1024, 1024); N = pri.shape[0]
    loc, conf = synth.detect_inputs(B, pri, 5050, 0.05)
    det = Detect(2, 0, 750, 0.05, 0.3)
    a = [torch.from_numpy(x).cuda() for x in (loc, conf, pri)]
    ms, ms_min = timed(lambda: det(*a))
    alg = B * (24 * N + 30000) + 16 * N
    print(json.dumps({"workload": "Detect B=64 @1024x1024 (N=87,360), conf 0.05, nms 0.3 (config 5 shard shape)", "ms": ms, "ms_min": ms_min,
                      "frames_per_s": B / (ms * 1e-3), "algorithmic_bytes": alg, "roofline_frac_of_measured_hbm": alg / (ms * 1e-3) / 1e9 / PEAK,
                      "candidates_per_image": float((conf[..., 1] > 0.05).sum(1).mean())}))


def detect512():
    """BASELINE config 5 on ONE GPU: B=512 @1024x1024 (N=87,360).  Inputs are generated on the device (1.07 GB)."""
    from fdt_b200.layers import Detect
    B = 512
    pri = synth.priors_numpy(1024, 1024); N = pri.shape[0]
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    loc = torch.randn((B, N, 4), device="cuda", generator=g) * 0.5
    d = torch.randn((B, N), device="cuda", generator=g) * 2.0 - 4.5
    s1 = torch.sigmoid(d)
    conf = torch.stack([1 - s1, s1], -1).contiguous()
    det = Detect(2, 0, 750, 0.05, 0.3)
    p = torch.from_numpy(pri).cuda()
    ms, ms_min = timed(lambda: det(loc, conf, p), reps=10)
    alg = B * (24 * N + 30000) + 16 * N
    print(json.dumps({"workload": "Detect B=512 @1024x1024 (N=87,360) on one GPU (config 5 unsharded; device-generated inputs, scores may tie)",
                      "ms": ms, "ms_min": ms_min, "frames_per_s": B / (ms * 1e-3), "algorithmic_bytes": alg,
                      "roofline_frac_of_measured_hbm": alg / (ms * 1e-3) / 1e9 / PEAK,
                      "candidates_per_image": float((conf[..., 1] > 0.05).sum(1).float().mean())}))


def priorbox():
    from fdt_b200.layers import PriorBoxLayer
    layer = PriorBoxLayer(640, 640)
    fm = synth.feature_maps(640, 640)
    ms, _ = timed(lambda: [layer(i, fw, fh) for i, (fw, fh) in enumerate(fm)])
    ref = orc.PriorBoxLayer(640, 640)
    cpu_s, _ = cpu_time(lambda: [ref(i, fw, fh) for i, (fw, fh) in enumerate(fm)], budget=2.0)
    print(json.dumps({"workload": "PriorBoxLayer(640,640) x 6 levels (34,125 priors)", "ms_6_launches_incl_python": ms, "cpu_port_ms": cpu_s * 1e3}))


if __name__ == "__main__":
    for w in sys.argv[1:] or ["multibox", "tracker", "detect1024", "priorbox"]:
        globals()[w]()

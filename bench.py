#!/usr/bin/env python
"""bench.py -- frames/s of Detect (decode + top-k + NMS) at 640x640, batch 64 per GPU (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode random|clustered]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one Detect pass over one batch of B=64 synthetic head outputs per GPU (weak scaling: every
rank owns its own 64 images; for N>1 the step ends with the gather of the fixed-shape detections block to rank 0
detections block, the only exchange the path has).  Prints ONE JSON line (rank 0).

  value     frames/s, inputs resident in HBM: the K steps are issued back to back over 3 rotating input sets (165 MB per
            GPU, larger than the 126 MB L2, so no step finds its inputs cached) and timed on the device with ONE CUDA-event
            pair on the launching stream, bracketed by barrier + synchronize; max over ranks.
  latency   the same step timed one at a time (own event pair, 256 MiB L2 flush + spin before it): what one isolated
            call costs; the gap to 1/value is launch latency that back-to-back issue hides.
  e2e       frames/s through the public API with pinned HOST tensors: Detect.__call__ -> fdt_detect_host
            (H2D copies + kernels + D2H of the detections inside the timed region, wall clock).
  roofline  dominant kernel (k_sort_nms) timed live with its own event pair via the stage entry points.
  cpu_baseline  the C oracle port (oracle/, the checker) on the host cores, bounded sample.
--impl reference times that CPU port with all host threads as the reference arm.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

B_PER_GPU = 64
WIDTH = HEIGHT = 640
TOP_K, NMS_TOP_K, CONF_T, NMS_T = 750, 5000, 0.05, 0.3
SEED = 20262
ROTATE = 3                                                # input sets per GPU (3 x 54.9 MB > L2)
METRIC = "frames/sec for Detect (decode+top-k+NMS) @640^2 batch 64; % HBM roofline"      # BASELINE.json metric, both arms


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="random", choices=["random", "clustered"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="peer", choices=["peer", "peer-barrier", "peer-all", "nccl"],
                    help="N>1: peer = rows stored into rank 0's gathered block by the NMS kernel over NVLink, completion signalled by the "
                         "kernel itself (fused gather); peer-barrier = the same followed by a symmetric-memory barrier; "
                         "peer-all = into every rank's block + barrier (fused all-gather); nccl = all_gather after the kernel")
    return ap.parse_args()


GATHER_NOTE = {"peer": ", gather to rank 0 fused into k_sort_nms (NVLink peer stores; completion signals published / awaited by the kernel "
                       "over symmetric memory: only rank 0 waits)",
               "peer-barrier": ", gather to rank 0 fused into k_sort_nms (NVLink peer stores + symmetric-memory barrier)",
               "peer-all": ", all-gather fused into k_sort_nms (NVLink peer stores into every rank's block + symmetric-memory barrier)",
               "nccl": ", NCCL all-gather of detections"}


def workload_config(mode, n_gpus, gather="nccl"):
    return {"workload": f"Detect batch {B_PER_GPU}/GPU @{WIDTH}x{HEIGHT} (34,125 priors), conf_thresh {CONF_T}, "
                        f"top_k {TOP_K}, nms_top_k {NMS_TOP_K}, NMS {NMS_T}; synthetic heads mode={mode} seed={SEED}",
            "global_batch": B_PER_GPU * n_gpus, "priors": 34125, "classes": 2,
            "parallelism": f"batch-sharded x{n_gpus}" + (GATHER_NOTE[gather] if n_gpus > 1 else "")}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_threads():
    """all hardware threads this process may use (torchrun exports OMP_NUM_THREADS=1, which must not throttle the CPU arm)"""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:                                             # noqa: BLE001
        return os.cpu_count() or 1


def cpu_oracle_rate(loc, conf, pri, budget_s, early_exit, threads=0):
    """frames/s of the C oracle port on the host cores; repeats the batch until ~budget_s of wall time."""
    from oracle import oracle as orc
    det = orc.Detect(2, 0, TOP_K, CONF_T, NMS_T)
    det.early_exit = early_exit
    threads = threads if threads > 0 else host_threads()
    det.n_threads = threads
    det(loc[:2], conf[:2], pri)                                   # page in
    n, t0 = 0, time.perf_counter()
    while True:
        det(loc, conf, pri)
        n += loc.shape[0]
        dt = time.perf_counter() - t0
        if dt >= budget_s or n >= 64 * 64:
            break
    return n / dt, n, dt, threads


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag, self.ok = [], set(), threading.Event(), False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:                                    # noqa: BLE001
            self.err = repr(e)
        self.active = False

    NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
             0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
             0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def sample(self):
        nv = self.nv
        mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:                                         # noqa: BLE001
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        if self.active:
            self.samples.append(mhz)
            for bit, name in self.NAMES.items():
                if r & bit and name != "gpu_idle":
                    self.reasons.add(name)

    def run(self):
        while self.ok and not self.stop_flag.is_set():
            try:
                self.sample()
            except Exception:                                     # noqa: BLE001
                pass
            time.sleep(0.0005)

    def result(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": getattr(self, "err", "nvml unavailable")}
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def run_reference(args):
    """Reference arm: the reference's CPU algorithm (C oracle port, run to completion like box_utils.nms)
    on the host cores with all threads; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from fdt_b200 import synth
    pri = synth.priors_numpy(WIDTH, HEIGHT)
    loc, conf = synth.detect_inputs(B_PER_GPU, pri, SEED, CONF_T, args.mode)
    from oracle import oracle as orc
    det = orc.Detect(2, 0, TOP_K, CONF_T, NMS_T)
    det.early_exit = False
    cores = host_threads()
    det.n_threads = cores
    for _ in range(max(args.warmup, 1)):
        det(loc, conf, pri)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        det(loc, conf, pri)
    dt = time.perf_counter() - t0
    fps = args.steps * B_PER_GPU / dt
    line = {"impl": "reference", "metric": METRIC, "value": fps,
            "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args.mode, 1),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} x the full B={B_PER_GPU} batch, C oracle (oracle/fdt_oracle.c), "
                                       f"OpenMP over images, NMS run to completion as the reference does"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import fdt_b200  # noqa: F401
    from fdt_b200 import _lib, synth
    from fdt_b200.layers import Detect

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.check(_lib.lib().fdt_device_check(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    pri_np = synth.priors_numpy(WIDTH, HEIGHT)
    N = pri_np.shape[0]
    # ROTATE input sets per GPU: 3 x 54.9 MB = 165 MB > 126 MB of L2, so back-to-back steps read their inputs from HBM
    sets = []
    for r in range(ROTATE):
        l_np, c_np = synth.detect_inputs(B_PER_GPU, pri_np, SEED + rank + 1000 * r, CONF_T, args.mode)
        if r == 0:
            loc_np, conf_np = l_np, c_np
        sets.append((torch.from_numpy(l_np).to(dev), torch.from_numpy(c_np).to(dev)))
    pri = torch.from_numpy(pri_np).to(dev)
    cur = [0]                                            # which input set the next step reads

    def inputs():
        return sets[cur[0] % ROTATE]
    loc, conf = sets[0]
    det = Detect(2, 0, TOP_K, CONF_T, NMS_T)
    B, C = B_PER_GPU, 2
    L = _lib.lib()
    out = torch.empty((B, C, TOP_K, 5), dtype=torch.float32, device=dev)
    gathered = torch.empty((world * B, C, TOP_K, 5), dtype=torch.float32, device=dev) if world > 1 else None
    ws = _lib.workspace(L.fdt_detect_workspace_bytes(B, N, C), dev, "bench")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)              # > 126 MB L2
    st = _lib.stream_ptr()

    def stage1():
        _lib.check(L.fdt_detect_threshold_compact(inputs()[1].data_ptr(), B, N, C, CONF_T, ws.data_ptr(), ws.numel(), st))

    def stage2():
        _lib.check(L.fdt_detect_sort_nms(inputs()[0].data_ptr(), pri.data_ptr(), B, N, C, TOP_K, NMS_TOP_K, NMS_T, 0.1, 0.2,
                                         out.data_ptr(), None, None, ws.data_ptr(), ws.numel(), st))

    gather = args.gather if world > 1 else "nccl"
    peer = None
    if world > 1 and gather in ("peer", "peer-barrier", "peer-all"):
        try:
            from fdt_b200.sharding import PeerGatherDetect
            peer = PeerGatherDetect(det, B, dest="all" if gather == "peer-all" else 0, signal="kernel" if gather == "peer" else "barrier")
        except Exception as e:                          # noqa: BLE001  (symmetric memory unavailable: keep the NCCL gather)
            if rank == 0:
                print(f"peer gather unavailable ({e!r}); using NCCL all_gather", file=sys.stderr)
            gather = "nccl"

    def stage2_gather():
        if peer is not None:
            hdl = peer.hdls[peer.turn]
            peer.turn ^= 1
            if peer.signal == "kernel":
                peer.epoch += 1
                _lib.check(L.fdt_detect_sort_nms_gather_signal(inputs()[0].data_ptr(), pri.data_ptr(), B, N, C, TOP_K, NMS_TOP_K, NMS_T, 0.1, 0.2,
                                                               int(hdl.buffer_ptrs_dev), int(peer.sig_hdl.buffer_ptrs_dev), world, rank, 0,
                                                               peer.epoch, rank * B, ws.data_ptr(), ws.numel(), st))
            else:
                ptrs, n_dst = peer.dest_ptrs(hdl)
                _lib.check(L.fdt_detect_sort_nms_peers(inputs()[0].data_ptr(), pri.data_ptr(), B, N, C, TOP_K, NMS_TOP_K, NMS_T, 0.1, 0.2,
                                                       ptrs, n_dst, rank * B, ws.data_ptr(), ws.numel(), st))
                hdl.barrier()
        else:
            stage2()
            if world > 1:
                dist.all_gather_into_tensor(gathered, out)

    def detect_call():
        lc = inputs()
        _lib.check(L.fdt_detect(lc[0].data_ptr(), lc[1].data_ptr(), pri.data_ptr(), B, N, C, TOP_K, NMS_TOP_K, CONF_T, NMS_T, 0.1, 0.2,
                                out.data_ptr(), None, None, ws.data_ptr(), ws.numel(), st))

    def step():
        # N=1: the public C-ABI call (K2 + K3 with programmatic dependent launch, nothing in between).
        # N>1: the same two kernels through the stage entry points so that K3 can store into the peers' blocks.
        if world == 1:
            detect_call()
        else:
            stage1()
            stage2_gather()
        cur[0] += 1

    spin_cycles = 250_000
    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        torch.cuda._sleep(spin_cycles)
        step()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    # ---- throughput: K steps back to back over the rotating input sets, one event pair, barrier + synchronize on both sides
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.active = True
    wall0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(2_000_000)                        # ~1 ms of GPU spin: the host gets ahead, the pair sees device time only
    e0.record()
    for i in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - wall0
    total_ms = float(e0.elapsed_time(e1))
    # ---- latency: the same step in isolation (own event pair; L2 flushed and the GPU spinning while the host enqueues it)
    n_lat = max(10, min(args.steps, 30))
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(n_lat)]
    for i in range(n_lat):
        flush.zero_()                                   # L2 flush, outside the event pairs
        torch.cuda._sleep(spin_cycles)
        ev[i][0].record()
        step()
        ev[i][1].record()
    torch.cuda.synchronize()
    if sampler.ok:
        try:
            sampler.sample()
        except Exception:                               # noqa: BLE001
            pass
    sampler.active = False
    sampler.stop_flag.set()
    step_ms = [e[0].elapsed_time(e[1]) for e in ev]
    k3_ms = None
    if world == 1:
        # dominant kernel alone: same launches through the stage entry points with an event pair around stage 2 only
        k3_ms = []
        for _ in range(max(10, min(args.steps, 30))):
            flush.zero_()
            torch.cuda._sleep(spin_cycles)
            stage1()
            a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            stage2()
            b2.record()
            torch.cuda.synchronize()
            k3_ms.append(a.elapsed_time(b2))
    lat_ms = float(np.mean(step_ms))
    if world > 1:
        t = torch.tensor([total_ms, lat_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, lat_ms = float(t[0].item()), float(t[1].item())
    ms_per_step = total_ms / args.steps
    value = world * B * args.steps / (total_ms * 1e-3)

    # ---- dominant kernel alone (k_sort_nms), timed live at N=1 with its own event pair
    peak, peak_src = peaks()
    roofline = None
    if world == 1:
        k3 = float(np.mean(k3_ms))
        k3_bytes = B * (16 * N + C * TOP_K * 5 * 4) + 16 * N           # loc + output rows per image, priors once
        step_bytes = B * (24 * N + C * TOP_K * 5 * 4) + 16 * N         # 849,000 B/image + priors
        roofline = {"bound": "hbm", "kernel": "k_sort_nms (select/sort + decode + lazy NMS + output rows)",
                    "achieved": k3_bytes / (k3 * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": k3_bytes / (k3 * 1e-3) / 1e9 / peak,
                    "traffic": 12.554e6,     # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full of this config
                    "traffic_source": "profiles/r01_sortnms_v6_phases.txt (12.547 MB read, 7.2 KB written)",
                    "peak_source": peak_src,
                    "kernel_ms": k3, "kernel_share_of_step": k3 / lat_ms,
                    "kernel_timing": "own event pair around the stage-2 launch of an isolated step, L2 flushed before the step "
                                     "(share = kernel_ms / latency.ms_per_step, both measured on isolated steps)",
                    "algorithmic_bytes_per_launch": k3_bytes,
                    "step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms_per_step * 1e-3) / 1e9,
                             "frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak},
                    "note": "NMS is O(K^2) IoU work on SM ALUs/LSU, not a streaming kernel; frac is bytes/time as the spec asks"}

    # ---- end to end through the public API with pinned host tensors
    loc_h, conf_h, pri_h = (torch.from_numpy(a).pin_memory() for a in (loc_np, conf_np, pri_np))
    for _ in range(5):
        o = det(loc_h, conf_h, pri_h)
    e2e_steps = max(20, min(args.steps, 100))
    if world > 1:
        dist.barrier()
    per = []
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        t1 = time.perf_counter()
        o = det(loc_h, conf_h, pri_h)                   # returns after the D2H of the detections completed
        per.append(time.perf_counter() - t1)
    e2e_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e = {"value": world * B * e2e_steps / e2e_dt, "unit": "frames/s",
           "h2d_bytes_per_step": int(conf_h.numel() * 4 + pri_h.numel() * 4),
           "h2d_note": "conf + priors are copied (cudaMemcpyAsync from pinned memory); the pinned loc tensor (%d bytes) is NOT copied: "
                       "k_sort_nms gathers only the rows NMS decodes (<= 1024 x 16 B per image and round) from host memory over PCIe" % int(loc_h.numel() * 4),
           "host_input_bytes_per_step": int(loc_h.numel() * 4 + conf_h.numel() * 4 + pri_h.numel() * 4),
           "d2h_bytes_per_step": int(o.numel() * 4), "steps": e2e_steps, "ms_per_step": 1e3 * e2e_dt / e2e_steps,
           "ms_per_step_median": 1e3 * statistics.median(per), "ms_per_step_max": 1e3 * max(per),
           "api": "fdt_b200.layers.Detect.__call__(pinned CPU tensors) -> fdt_detect_host"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        fps, n, dt, cores = cpu_oracle_rate(loc_np, conf_np, pri_np, budget_s=10.0, early_exit=False)
        cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"{n} images ({n // B} x the B={B} batch) in {dt:.1f} s, C oracle port with OpenMP over images, "
                         f"NMS run to completion as box_utils.nms does (the reference itself is a Python loop, ~1-2 frames/s)"}

    line = {"metric": METRIC,
            "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": dict(workload_config(args.mode, world, gather),
                                                                l2=f"inputs larger than L2: {ROTATE} rotating input sets per GPU ({ROTATE * 54.9:.0f} MB > 126 MB), steps issued back to back, "
                                                                   "one CUDA-event pair around the K steps"),
            "latency": {"ms_per_step": lat_ms, "frames_per_s": world * B / (lat_ms * 1e-3), "steps": n_lat,
                        "how": "each step alone with its own event pair; 256 MiB L2 flush + 0.1 ms GPU spin before it (outside the pair)"},
            "clocks": sampler.result(), "e2e": e2e, "gpu_launches": 3 * args.steps,      # k_zero_counters, k_threshold_compact, k_sort_nms
            "wall_s_timed_region": wall}
    if roofline:
        line["roofline"] = roofline
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

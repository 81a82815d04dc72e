"""Diagnostics (torchrun, >= 2 GPUs): what the fused multi-GPU gather adds to the Detect step (headline shape, 64 images per rank):
local output / rows stored into every rank's block / + symmetric-memory barrier."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import fdt_b200
from fdt_b200 import _lib, synth
from fdt_b200.layers import Detect
from fdt_b200.sharding import PeerGatherDetect

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, C, TOP_K = 64, 2, 750
pri_np = synth.priors_numpy(640, 640)
loc_np, conf_np = synth.detect_inputs(B, pri_np, 20262 + rank, 0.05, "random")
N = pri_np.shape[0]
loc, conf, pri = (torch.from_numpy(a).to(dev) for a in (loc_np, conf_np, pri_np))
det = Detect(2, 0, TOP_K, 0.05, 0.3)
peer = PeerGatherDetect(det, B)
L = _lib.lib()
out = torch.empty((B, C, TOP_K, 5), device=dev)
ws = _lib.workspace(L.fdt_detect_workspace_bytes(B, N, C), dev, "diag")
st = _lib.stream_ptr()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def k2():
    _lib.check(L.fdt_detect_threshold_compact(conf.data_ptr(), B, N, C, 0.05, ws.data_ptr(), ws.numel(), st))


def local_step():
    k2()
    _lib.check(L.fdt_detect_sort_nms(loc.data_ptr(), pri.data_ptr(), B, N, C, TOP_K, 5000, 0.3, 0.1, 0.2, out.data_ptr(), None, None,
                                     ws.data_ptr(), ws.numel(), st))


def peers(barrier):
    def f():
        hdl = peer.hdls[peer.turn]
        peer.turn ^= 1
        k2()
        _lib.check(L.fdt_detect_sort_nms_peers(loc.data_ptr(), pri.data_ptr(), B, N, C, TOP_K, 5000, 0.3, 0.1, 0.2,
                                               int(hdl.buffer_ptrs_dev), world, rank * B, ws.data_ptr(), ws.numel(), st))
        if barrier:
            hdl.barrier()
    return f


peer_sig = PeerGatherDetect(det, B, dest=0, signal="kernel")
peer_root = PeerGatherDetect(det, B, dest=0)


def root_barrier():
    hdl = peer_root.hdls[peer_root.turn]
    peer_root.turn ^= 1
    k2()
    ptrs, n_dst = peer_root.dest_ptrs(hdl)
    _lib.check(L.fdt_detect_sort_nms_peers(loc.data_ptr(), pri.data_ptr(), B, N, C, TOP_K, 5000, 0.3, 0.1, 0.2, ptrs, n_dst, rank * B,
                                           ws.data_ptr(), ws.numel(), st))
    hdl.barrier()


def root_signal():
    hdl = peer_sig.hdls[peer_sig.turn]
    peer_sig.turn ^= 1
    peer_sig.epoch += 1
    k2()
    _lib.check(L.fdt_detect_sort_nms_gather_signal(loc.data_ptr(), pri.data_ptr(), B, N, C, TOP_K, 5000, 0.3, 0.1, 0.2,
                                                   int(hdl.buffer_ptrs_dev), int(peer_sig.sig_hdl.buffer_ptrs_dev), world, rank, 0, peer_sig.epoch,
                                                   rank * B, ws.data_ptr(), ws.numel(), st))


def barrier_only():
    peer.hdls[0].barrier()


def timed(fn, reps=30):
    ts = []
    for i in range(reps + 5):
        flush.zero_()
        dist.barrier()
        torch.cuda._sleep(400_000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if i >= 5:
            ts.append(a.elapsed_time(b) * 1e3)
    t = torch.tensor([float(np.mean(ts))], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for name, fn in (("local output", local_step), ("peer stores, no barrier", peers(False)), ("peer stores + barrier", peers(True)),
                 ("gather to rank 0 + barrier", root_barrier), ("gather to rank 0, kernel signal", root_signal),
                 ("barrier alone", barrier_only)):
    v = timed(fn)
    if rank == 0:
        print(f"{name:28s} {v:7.1f} us (max over {world} ranks)")
dist.destroy_process_group()

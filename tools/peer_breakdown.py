"""Diagnostics (torchrun, >= 2 GPUs): what the multi-GPU gather adds to the Detect step (headline shape, 64 images per rank), as the
step is timed by bench.py -- K calls back to back per variant, max over ranks -- and per rank, so that the limiter shows:

  local            every rank runs fdt_detect into its own output (no exchange)
  gather signalled rows stored into rank 0's gathered block by the NMS kernel, completion signals through symmetric memory
  all-gather sig.  rows stored into every rank's block, signalled
  gather + barrier rows stored into rank 0's block, symmetric-memory barrier after every call
  nccl             local output + NCCL all_gather_into_tensor

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/peer_breakdown.py [K]
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import fdt_b200  # noqa: F401
from fdt_b200 import _lib, synth
from fdt_b200.layers import Detect
from fdt_b200.sharding import PeerGatherDetect

K = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, C, TOP_K = 64, 2, 750
pri_np = synth.priors_numpy(640, 640)
N = pri_np.shape[0]
sets = [tuple(torch.from_numpy(a).to(dev) for a in synth.detect_inputs(B, pri_np, 20262 + rank + 1000 * r, 0.05)) for r in range(3)]
pri = torch.from_numpy(pri_np).to(dev)
det = Detect(2, 0, TOP_K, 0.05, 0.3)
L = _lib.lib()
st = _lib.stream_ptr()
ws = torch.empty(L.fdt_detect_workspace_bytes_depth(B, N, C, 4), dtype=torch.uint8, device=dev)
outs = [torch.empty((B, C, TOP_K, 5), device=dev) for _ in range(4)]
gathered = [torch.empty((world * B, C, TOP_K, 5), device=dev) for _ in range(4)]
step_no = [0]


def args12():
    l, c = sets[step_no[0] % 3]
    return (l.data_ptr(), c.data_ptr(), pri.data_ptr(), B, N, C, TOP_K, 5000, 0.05, 0.3, 0.1, 0.2)


def local_step():
    _lib.check(L.fdt_detect(*args12(), outs[step_no[0] % 4].data_ptr(), None, None, ws.data_ptr(), ws.numel(), st))


def make_peer(dest, signal):
    peer = PeerGatherDetect(det, B, dest=dest, signal=signal)

    def f():
        hdl = peer.hdls[peer.turn]
        peer.turn = (peer.turn + 1) % peer.RING
        ptrs, n_dst = peer.dest_ptrs(hdl)
        if signal.startswith("kernel"):
            peer.epoch += 1
            root = -1 if dest == "all" else int(dest)
            call = L.fdt_detect_gather_store if peer.await_stream is not None else L.fdt_detect_gather_signal
            _lib.check(call(*args12(), ptrs, n_dst, int(peer.sig_hdl.buffer_ptrs_dev), world, rank, root, peer.epoch, peer.RING, rank * B,
                            ws.data_ptr(), ws.numel(), st))
            peer._last_ws = ws
        else:
            _lib.check(L.fdt_detect_peers(*args12(), ptrs, n_dst, rank * B, ws.data_ptr(), ws.numel(), st))
            hdl.barrier()
    f.peer = peer
    return f


def nccl_step():
    o = outs[step_no[0] % 4]
    _lib.check(L.fdt_detect(*args12(), o.data_ptr(), None, None, ws.data_ptr(), ws.numel(), st))
    dist.all_gather_into_tensor(gathered[step_no[0] % 4], o)


def timed(fn):
    for _ in range(5):
        fn(); step_no[0] += 1
    res = []
    for _ in range(5):
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(2_000_000)
        a.record()
        for _ in range(K):
            fn(); step_no[0] += 1
        if getattr(fn, "peer", None) is not None:
            fn.peer.wait_ready()
        b.record()
        torch.cuda.synchronize()
        dist.barrier()
        res.append(a.elapsed_time(b) * 1e3 / K)
    mine = sorted(res)[len(res) // 2]
    t = torch.tensor([mine], dtype=torch.float64, device=dev)
    allr = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allr, t)
    return [float(x.item()) for x in allr]


variants = [("local (no exchange)", local_step), ("gather to rank 0, signalled", make_peer(0, "kernel")),
            ("  .. awaits on their own stream", make_peer(0, "kernel-side")),
            ("all-gather, signalled", make_peer("all", "kernel")), ("gather to rank 0 + barrier", make_peer(0, "barrier")),
            ("nccl all-gather", nccl_step)]
if os.environ.get("FDT_BREAKDOWN_SHORT"):
    variants = variants[:3]
for name, fn in variants:
    per_rank = timed(fn)
    if rank == 0:
        print(f"{name:30s} max {max(per_rank):6.2f} us/step   per rank: " + " ".join(f"{v:6.2f}" for v in per_rank), flush=True)
dist.destroy_process_group()

"""Multi-GPU check (not collected by pytest; run with torchrun on >= 2 GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_check.py
Every rank detects its shard; the fused peer-memory gather and the NCCL all-gather must both reproduce, byte for byte, what
one process computes on the concatenated batch (checked against the CPU oracle on rank 0)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fdt_b200  # noqa: E402
from fdt_b200 import synth  # noqa: E402
from fdt_b200.layers import Detect  # noqa: E402
from fdt_b200.sharding import PeerGatherDetect, ShardedDetect, shard_range  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    b_local = 8
    B = b_local * world
    pri = synth.priors_numpy(640, 640)
    loc, conf = synth.detect_inputs(B, pri, 777, 0.05)
    lo, hi = shard_range(B, rank, world)
    cu = lambda a: torch.from_numpy(a).cuda()
    det = Detect(2, 0, 750, 0.05, 0.3)
    nccl = ShardedDetect(det, gather="block")(cu(loc[lo:hi]), cu(conf[lo:hi]), cu(pri)).cpu().numpy()
    packed = ShardedDetect(det, gather="packed")(cu(loc[lo:hi]), cu(conf[lo:hi]), cu(pri)).cpu().numpy()
    outs = []
    peer = PeerGatherDetect(det, b_local)
    outs += [peer(cu(loc[lo:hi]), cu(conf[lo:hi]), cu(pri)).cpu().numpy() for _ in range(5)]
    # different images in the same blocks: stale rows of the previous call must be overwritten everywhere
    sh = (np.arange(B) + 3) % B
    o2 = peer(cu(loc[sh][lo:hi]), cu(conf[sh][lo:hi]), cu(pri)).cpu().numpy()
    outs.append(o2[np.argsort(sh)])
    ok = nccl.tobytes() == packed.tobytes() and all(o.tobytes() == nccl.tobytes() for o in outs)
    # gather to a root: only the root's block is written; with the barrier after the kernel, or signalled by the kernel itself
    for root, signal in ((0, "barrier"), (world - 1, "barrier"), (0, "kernel"), (world - 1, "kernel")):
        pr = PeerGatherDetect(det, b_local, dest=root, signal=signal)
        for it in range(6):
            sh = (np.arange(B) + it) % B           # a different batch every call: stale rows must never survive
            o = pr(cu(loc[sh][lo:hi]), cu(conf[sh][lo:hi]), cu(pri))
            if rank == root:
                torch.cuda.synchronize()
                ok = ok and o.cpu().numpy()[np.argsort(sh)].tobytes() == nccl.tobytes()
    if rank == 0:
        from oracle import oracle as orc
        o = orc.Detect(2, 0, 750, 0.05, 0.3); o.early_exit = True
        ok = ok and nccl.tobytes() == o(loc, conf, pri).tobytes()
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU_CHECK", "PASS" if int(t.item()) == 1 else "FAIL", "world", world)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()

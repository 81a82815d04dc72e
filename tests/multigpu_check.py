"""Multi-GPU check (not collected by pytest; run with torchrun on >= 2 GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_check.py
Every rank detects its shard; the fused peer-memory gather (barrier / signalled, to a root / to every rank) and the NCCL
all-gather must all reproduce, byte for byte, what one process computes on the concatenated batch (checked against the CPU
oracle on rank 0); the sharded MultiBoxLoss and its gradients must equal the single-process loss on the whole batch."""
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


def orc_loss(loc, conf, pri, targets):
    from oracle import oracle as orc
    return orc.multibox_loss(loc, conf, pri, targets, 0.35, 3, True, want_aux=False)      # bipartite=True: MultiBoxLoss' default (multibox_loss.py:33)


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
    # a second batch with FEW detections per image (clustered faces): the signalled gather only ships the rows that exist and
    # clears what an earlier epoch left behind them in the same ring block
    loc_c, conf_c = synth.detect_inputs(B, pri, 778, 0.05, "clustered")
    nccl_c = ShardedDetect(det, gather="block")(cu(loc_c[lo:hi]), cu(conf_c[lo:hi]), cu(pri)).cpu().numpy()
    packed = ShardedDetect(det, gather="packed")(cu(loc[lo:hi]), cu(conf[lo:hi]), cu(pri)).cpu().numpy()
    outs = []
    peer = PeerGatherDetect(det, b_local)
    outs += [peer(cu(loc[lo:hi]), cu(conf[lo:hi]), cu(pri)).cpu().numpy() for _ in range(5)]
    # different images in the same blocks: stale rows of the previous call must be overwritten everywhere
    sh = (np.arange(B) + 3) % B
    o2 = peer(cu(loc[sh][lo:hi]), cu(conf[sh][lo:hi]), cu(pri)).cpu().numpy()
    outs.append(o2[np.argsort(sh)])
    fails = []

    def check(name, cond):
        if not cond:
            fails.append(name)
            print(f"rank {rank}: CHECK FAILED: {name}", flush=True)
        return cond
    ok = check("packed == block", nccl.tobytes() == packed.tobytes())
    ok = check("peer all-gather (barrier) == nccl", all(o.tobytes() == nccl.tobytes() for o in outs)) and ok
    # gather to a root: only the root's block is written; with the barrier after the kernel, or signalled through symmetric memory
    # (10 calls back to back: more than the ring of gathered blocks and than the workspace ring, a different batch every call so
    # that stale rows can never pass)
    for root, signal in ((0, "barrier"), (world - 1, "barrier"), (0, "kernel"), (world - 1, "kernel"), ("all", "kernel"),
                         (0, "kernel-side"), (world - 1, "kernel-side"), ("all", "kernel-side")):
        pr = PeerGatherDetect(det, b_local, dest=root, signal=signal)
        got = []
        for it in range(12):
            sh = (np.arange(B) + it) % B
            few = it % 3 == 1 or it in (6, 7, 8)                               # full -> few -> full ... and few three times in a row
            l_, c_ = (loc_c, conf_c) if few else (loc, conf)
            o = pr(cu(l_[sh][lo:hi]), cu(c_[sh][lo:hi]), cu(pri))
            pr.wait_ready()                    # (kernel-side: the await kernels run on their own stream; a no-op otherwise)
            got.append((sh, few, o.clone()))   # an ordinary kernel behind the call: on a destination it must see every rank's rows
        torch.cuda.synchronize()
        pr.check()
        if root == "all" or rank == root:
            for it, (sh, few, o) in enumerate(got):
                want = nccl_c if few else nccl
                ok = check(f"gather dest={root} signal={signal} call {it}", o.cpu().numpy()[np.argsort(sh)].tobytes() == want.tobytes()) and ok
        dist.barrier()
    # MultiBoxLoss sharded over the ranks (multibox_loss.py:130-135: one global normalisation) == one process on the whole batch
    from fdt_b200.layers import MultiBoxLoss
    from fdt_b200.sharding import sharded_multibox_loss
    l2, c2, targets = synth.multibox_inputs(B, pri, 98, 0, 40)
    crit = MultiBoxLoss(2, 0.35, True, 0, True, 3, 0.35, False)
    la = cu(l2[lo:hi]).requires_grad_(True); ca = cu(c2[lo:hi]).requires_grad_(True)
    ll, lc = sharded_multibox_loss(crit, (la, ca, cu(pri)), [cu(t) for t in targets[lo:hi]])
    (ll + lc).backward()
    lf = cu(l2).requires_grad_(True); cf = cu(c2).requires_grad_(True)
    fl, fc = MultiBoxLoss(2, 0.35, True, 0, True, 3, 0.35, False)((lf, cf, cu(pri)), [cu(t) for t in targets])
    (fl + fc).backward()
    ok = check("sharded loss == single-process loss", np.allclose([float(ll.detach()), float(lc.detach())], [float(fl.detach()), float(fc.detach())], rtol=1e-5)) and ok
    ok = check("sharded grad_loc", np.allclose(la.grad.cpu().numpy(), lf.grad[lo:hi].cpu().numpy(), rtol=1e-4, atol=1e-7)) and ok
    ok = check("sharded grad_conf", np.allclose(ca.grad.cpu().numpy(), cf.grad[lo:hi].cpu().numpy(), rtol=1e-4, atol=1e-7)) and ok
    if rank == 0:
        r = orc_loss(l2, c2, pri, targets)
        same = np.allclose([float(ll.detach()), float(lc.detach())], [r["loss_l"], r["loss_c"]], rtol=1e-5)
        if not same:
            print("sharded", float(ll.detach()), float(lc.detach()), "single", float(fl.detach()), float(fc.detach()), "oracle", r["loss_l"], r["loss_c"], flush=True)
        ok = check("sharded loss == oracle", same) and ok
    if rank == 0:
        from oracle import oracle as orc
        o = orc.Detect(2, 0, 750, 0.05, 0.3); o.early_exit = True
        ok = check("nccl block == oracle", nccl.tobytes() == o(loc, conf, pri).tobytes()) and ok
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU_CHECK", "PASS" if int(t.item()) == 1 else "FAIL", "world", world)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()

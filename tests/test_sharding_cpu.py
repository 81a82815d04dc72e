"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: contiguous batch sharding, the detections
all-gather (block and packed) and the MultiBoxLoss all-reduce.  The per-rank compute is stubbed with the CPU
oracle (tests may use it as the checker); the NCCL path runs the same code on GPUs."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_shard_range_partitions_exactly():
    from fdt_b200.sharding import shard_range
    for n in (0, 1, 7, 64, 512, 513):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


class OracleDetect:
    def __init__(self):
        from oracle import oracle as orc
        self.det = orc.Detect(2, 0, 50, 0.05, 0.3)
        self.det.n_threads = 1

    def __call__(self, loc, conf, pri, return_aux=False):
        out, counts, kept = self.det(loc.numpy(), conf.numpy(), pri.numpy(), return_aux=True)
        if return_aux:
            return torch.from_numpy(out), torch.from_numpy(counts), torch.from_numpy(kept)
        return torch.from_numpy(out)


class OracleLoss:
    def __call__(self, predictions, targets):
        from oracle import oracle as orc
        loc, conf, pri = predictions
        r = orc.multibox_loss(loc.numpy(), conf.numpy(), pri.numpy(), [t.numpy() for t in targets], 0.35, 3, False, n_threads=1)
        self.last_aux = (torch.from_numpy(r["loc_t"]), torch.from_numpy(r["conf_t"]), None)
        return torch.tensor(r["loss_l"]), torch.tensor(r["loss_c"])


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fdt_b200 import synth
        from fdt_b200.sharding import ShardedDetect, shard_range, sharded_multibox_loss
        pri = synth.priors_numpy(160, 160)
        loc, conf = synth.detect_inputs(B, pri, 99, 0.05)
        lo, hi = shard_range(B, rank, world)
        t = torch.from_numpy
        res = {}
        for mode in ("block", "packed"):
            sd = ShardedDetect(OracleDetect(), gather=mode)
            res[mode] = sd(t(loc[lo:hi]), t(conf[lo:hi]), t(pri)).numpy()
        l2, c2, targets = synth.multibox_inputs(B, pri, 98, 0, 6)
        ll, lc = sharded_multibox_loss(OracleLoss(), (t(l2[lo:hi]), t(c2[lo:hi]), t(pri)), [t(x) for x in targets[lo:hi]])
        res["loss"] = (float(ll), float(lc))
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [4, 5])
def test_sharded_equals_single_process_world2(B):
    from fdt_b200 import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + B) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    pri = synth.priors_numpy(160, 160)
    loc, conf = synth.detect_inputs(B, pri, 99, 0.05)
    single = OracleDetect()(torch.from_numpy(loc), torch.from_numpy(conf), torch.from_numpy(pri)).numpy()
    l2, c2, targets = synth.multibox_inputs(B, pri, 98, 0, 6)
    sl, sc = OracleLoss()((torch.from_numpy(l2), torch.from_numpy(c2), torch.from_numpy(pri)), [torch.from_numpy(x) for x in targets])
    for rank in (0, 1):
        assert got[rank]["block"].tobytes() == single.tobytes()          # byte-identical to the single-process output
        assert got[rank]["packed"].tobytes() == single.tobytes()
        np.testing.assert_allclose(got[rank]["loss"], (float(sl), float(sc)), rtol=1e-6)


def test_tracker_pack_frames_host_logic():
    """pack_frames (host side of the tracker drop-in): float32 rows, the float64 dummy row, empty frames and an empty video."""
    import numpy as np
    from fdt_b200 import tracker
    frames = [np.array([[1, 2, 3, 4, 0.5], [5, 6, 7, 8, 0.9]], np.float32), np.zeros((0, 5), np.float32),
              np.array([[0, 0, 0, 0, 0.4]]), [[9.5, 1, 2, 3, 0.7]]]
    dets, off = tracker.pack_frames(frames)
    assert dets.dtype == np.float64 and off.dtype == np.int64
    assert off.tolist() == [0, 2, 2, 3, 4]
    assert np.array_equal(dets, np.array([[1, 2, 3, 4, np.float32(0.5)], [5, 6, 7, 8, np.float32(0.9)], [0, 0, 0, 0, 0.4], [9.5, 1, 2, 3, 0.7]]))
    dets, off = tracker.pack_frames([])
    assert off.tolist() == [0] and dets.shape == (1, 5)
    dets, off = tracker.pack_frames([np.zeros((0, 5))])
    assert off.tolist() == [0, 0] and dets.shape == (1, 5)

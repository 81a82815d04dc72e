"""Multi-GPU use of the box pipeline: images are independent (detection.py:52 and multibox_loss.py:69 loop
over the batch with no cross-image state), so the batch dimension is split contiguously across ranks, one
process per GPU, with NO collective on the data path.  The only exchanges are

  * Detect: one all-gather of the detections -- either the fixed-shape [B_local, C, top_k, 5] block
    (byte-identical to the single-GPU output once concatenated) or counts + packed variable-length rows;
  * MultiBoxLoss: one all-reduce(sum) of (sum loss_l, sum loss_c, num_pos) before the division at
    multibox_loss.py:134-135.

The tracker is a serial chain per video: replicas only (one video per rank), no collective.
Backends: NCCL over NVLink on GPUs; the same code runs under gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous balanced split: rank r owns [lo, hi); the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world(group):
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def all_gather_ragged(t: torch.Tensor, group=None):
    """all-gather along dim 0 of tensors whose dim-0 length differs per rank (other dims equal)."""
    rank, world = _world(group)
    if world == 1:
        return t
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    if all(s == m for s in sizes):
        out = torch.empty((world * m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
        return out
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)


class ShardedDetect:
    """detect: a Detect-like callable (loc, conf, priors) -> [B_local, C, top_k, 5] run on this rank's images.

    __call__ returns the detections of ALL ranks' images in rank order (== the single-process result on the
    concatenated batch).  gather="block" exchanges the zero-padded block; gather="packed" exchanges
    counts[B, C] + only the rows that hold detections and rebuilds the block locally."""

    def __init__(self, detect, group=None, gather="block"):
        assert gather in ("block", "packed")
        self.detect, self.group, self.gather = detect, group, gather

    def __call__(self, loc_local, conf_local, priors):
        rank, world = _world(self.group)
        if world == 1 or self.gather == "block":
            out = self.detect(loc_local, conf_local, priors)
            return out if world == 1 else all_gather_ragged(out, self.group)
        # packed: the kernel's own row counts (rows are contiguous from 0 in every plane) -- no assumption about the sign of the
        # scores, no extra pass over the block
        try:
            out, counts, _ = self.detect(loc_local, conf_local, priors, return_aux=True)
        except TypeError:                            # a Detect-like callable without aux outputs: count the leading non-zero rows
            out = self.detect(loc_local, conf_local, priors)
            counts = (out != 0).any(-1).sum(-1).to(torch.int32)
        B, C, K, _ = out.shape
        idx = torch.arange(K, device=out.device).expand(B, C, K) < counts.unsqueeze(-1)
        rows = out[idx]                              # [sum counts, 5] in (image, class, rank) order
        counts_all = all_gather_ragged(counts, self.group)
        rows_all = all_gather_ragged(rows, self.group)
        full = torch.zeros((counts_all.shape[0], C, K, 5), dtype=out.dtype, device=out.device)
        idx_all = torch.arange(K, device=out.device).expand(counts_all.shape[0], C, K) < counts_all.unsqueeze(-1)
        full[idx_all] = rows_all
        return full


def sharded_multibox_loss(criterion, predictions_local, targets_local, group=None):
    """criterion: a MultiBoxLoss-like module whose forward sets .last_aux = (loc_t, conf_t, sel) and returns
    (loss_l, loss_c) normalised by the LOCAL number of positives.  Returns the losses normalised by the GLOBAL
    number of positives, exactly what one process would return on the concatenated batch."""
    loss_l, loss_c = criterion(predictions_local, targets_local)
    rank, world = _world(group)
    if world == 1:
        return loss_l, loss_c
    conf_t = criterion.last_aux[1]
    num_pos = (conf_t > 0).sum().to(torch.float32)
    b_local = torch.tensor(float(conf_t.shape[0]), device=num_pos.device)
    n_local = torch.where(num_pos > 0, num_pos, b_local)                    # multibox_loss.py:130-133
    dev = num_pos.device
    sums = torch.stack([loss_l.detach().to(dev) * n_local, loss_c.detach().to(dev) * n_local, num_pos, b_local])
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    n_glob = torch.where(sums[2] > 0, sums[2], sums[3])
    # keep the autograd graph of the local terms: d(global)/d(local inputs) = local grad * n_local / n_glob
    scale = (n_local / n_glob).to(loss_l.device)
    others_l = (sums[0] / n_glob).to(loss_l.device) - loss_l.detach() * scale
    others_c = (sums[1] / n_glob).to(loss_c.device) - loss_c.detach() * scale
    return loss_l * scale + others_l, loss_c * scale + others_c


class PeerGatherDetect:
    """Detect with the gather FUSED into the NMS kernel: k_sort_nms stores the detection rows of this rank's images straight
    into the destination ranks' gathered block [world * B_local, C, top_k, 5] over NVLink peer memory with 16-byte vector stores
    (torch symmetric memory provides the peer pointers); no NCCL collective is launched.  RING blocks alternate between calls so a
    rank that runs ahead never overwrites rows a slower peer is still reading.  The returned tensor is this rank's copy of the
    gathered block and stays valid until RING - 1 further calls have been made."""

    RING = 4

    def __init__(self, detect, b_local, group=None, dest="all", signal="barrier"):
        """dest="all": every rank ends up with the whole gathered block (all-gather).  dest=<rank>: only that rank does (gather
        to a root): each rank's rows cross NVLink once instead of world - 1 times; only the root's returned block is meaningful.
        signal="barrier": a symmetric-memory barrier follows the kernel (CUDA-graph capturable).
        signal="kernel": the completion signals travel through symmetric memory inside the call (fdt_detect_gather_signal): a source
        rank never waits for anybody (except for a destination that is RING calls behind), a destination enqueues a one-block
        await kernel behind its own NMS kernel -- consumers of the block follow it in stream order, the next call does not.
        signal="kernel-side": the same stores and signals, but no await kernel in the calls' stream (there it is a third grid per
        call in the window of grids the stream runs ahead by, and a resident one-block kernel keeps a whole SM from taking an NMS
        CTA: measured +1.3 us per call on the destination).  The await kernel depends on nothing but the signal array, so a
        destination enqueues it on a stream of its own when the rows are NEEDED: `wait_ready()` awaits the latest call -- epochs
        are monotonic, it covers every earlier call -- and makes the consumer's stream wait for it (an event).  Consumers that run
        in the calls' stream keep the ring's overwrite protection (a source only reuses a block once the destination has begun
        the call RING - 1 later, which in that stream follows the consumer)."""
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        assert signal in ("barrier", "kernel", "kernel-side")
        self.await_stream = None
        if signal == "kernel-side":
            signal = "kernel"
            self.await_stream = torch.cuda.Stream()
        self.dest, self.signal = dest, signal
        self._lib = _lib
        self.detect = detect
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.b_local = int(b_local)
        dev = torch.device("cuda", torch.cuda.current_device())
        shape = (self.world * self.b_local, detect.num_classes, detect.top_k, 5)
        self.bufs, self.hdls = [], []
        for _ in range(self.RING):
            t = symm_mem.empty(shape, dtype=torch.float32, device=dev)
            t.zero_()                       # the kernel never writes the background planes: zero once, zero forever
            self.hdls.append(symm_mem.rendezvous(t, self.group))
            self.bufs.append(t)
        self.sig = self.sig_hdl = None
        self.epoch = 0
        if signal == "kernel":
            self.sig = symm_mem.empty((max(2 * self.world, 64),), dtype=torch.int32, device=dev)    # uint32 epoch slots
            self.sig.zero_()
            self.sig_hdl = symm_mem.rendezvous(self.sig, self.group)
        self._workspaces = _lib.DetectWorkspaces()
        self._last_ws = None
        torch.cuda.synchronize()
        dist.barrier(self.group)            # every rank's blocks are zeroed before any peer stores rows into them
        self.turn = 0

    def __call__(self, loc, conf, priors):
        _lib = self._lib
        L = _lib.lib()
        d = self.detect
        B, N = loc.shape[0], priors.shape[0]
        assert B == self.b_local, "PeerGatherDetect was built for a fixed per-rank batch"
        dev = loc.device
        loc, conf, priors = _lib.dev_f32(loc, dev), _lib.dev_f32(conf, dev), _lib.dev_f32(priors, dev)
        ws = self._workspaces.get(B, N, d.num_classes, dev)
        st = _lib.stream_ptr()
        hdl, buf = self.hdls[self.turn], self.bufs[self.turn]
        self.turn = (self.turn + 1) % self.RING
        args = (loc.data_ptr(), conf.data_ptr(), priors.data_ptr(), B, N, d.num_classes, int(d.top_k), int(d.nms_top_k),
                float(d.conf_thresh), float(d.nms_thresh), float(d.variance[0]), float(d.variance[1]))
        ptrs, n_dst = self.dest_ptrs(hdl)
        if self.signal == "kernel":
            self.epoch += 1
            root = -1 if self.dest == "all" else int(self.dest)
            call = L.fdt_detect_gather_store if self.await_stream is not None else L.fdt_detect_gather_signal
            _lib.check(call(*args, ptrs, n_dst, int(self.sig_hdl.buffer_ptrs_dev), self.world, self.rank, root, self.epoch, self.RING,
                            self.rank * B, ws.data_ptr(), ws.numel(), st))
            self._last_ws = ws
            return buf           # on a destination: complete in stream order behind the await kernel (kernel-side: after wait_ready())
        _lib.check(L.fdt_detect_peers(*args, ptrs, n_dst, self.rank * B, ws.data_ptr(), ws.numel(), st))
        hdl.barrier()            # all ranks' rows have landed in the destination block(s)
        return buf

    def wait_ready(self, stream=None, ws=None):
        """kernel-side: makes `stream` (default: the current one) wait until the gathered blocks of all calls made so far are
        complete on this rank (one await kernel for the latest epoch on the await stream + an event).  A no-op for the other
        signal modes (stream order already covers them) and on a rank that is not a destination."""
        if self.await_stream is None or self.epoch == 0 or not (self.dest == "all" or int(self.dest) == self.rank):
            return
        _lib = self._lib
        ws = ws if ws is not None else self._last_ws
        _lib.check(_lib.lib().fdt_detect_gather_await(int(self.sig_hdl.buffer_ptrs_dev), self.world, self.rank, self.epoch,
                                                      ws.data_ptr(), self.await_stream.cuda_stream))
        ev = torch.cuda.Event()
        ev.record(self.await_stream)
        (stream if stream is not None else torch.cuda.current_stream()).wait_event(ev)

    def dest_ptrs(self, hdl):
        """(device array of destination block pointers, how many)."""
        if self.dest == "all":
            return int(hdl.buffer_ptrs_dev), self.world
        return int(hdl.buffer_ptrs_dev) + 8 * int(self.dest), 1          # one entry of the symmetric pointer table: the root's block

    def check(self):
        """Raises if a cross-rank wait timed out (a peer died or never made the call); synchronises."""
        bits = self._workspaces.status()
        if bits:
            raise RuntimeError(f"fdt_b200: fused gather timed out (status {bits:#x}: "
                               f"{'peer never signalled / acknowledged' if bits & 2 else 'local call never completed'})")

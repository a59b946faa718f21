"""Image sharding across the GPUs of one box and the single collective of the path.

Every (image, task) segment is independent in decode and NMS (reference
utils/general.py:424 loops over images; tasks are separate calls at
cerberusdet_inference.py:125-135), so a batch shards by image with no data-path
collective.  The only exchange is one gather of the padded detections
``[T, B_loc, max_det, 6]`` + counts ``[T, B_loc]`` to rank 0 (NCCL over NVLink on
GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_images: int, rank: int, world: int) -> range:
    """Contiguous block of images owned by ``rank``; remainder images go to the low ranks."""
    base, rem = divmod(n_images, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def shard_sizes(n_images: int, world: int) -> List[int]:
    return [len(shard_range(n_images, r, world)) for r in range(world)]


def gather_detections(
    dets: torch.Tensor, counts: torch.Tensor, dst: int = 0, n_images: Optional[int] = None, group=None
) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Gather per-rank ``dets[T, B_loc, max_det, 6]`` / ``counts[T, B_loc]`` to ``dst``.

    Returns ``(dets[T, B, max_det, 6], counts[T, B])`` on ``dst`` (image order = rank
    order = global image order) and ``(None, None)`` elsewhere.  With ``n_images`` given,
    ranks may hold unequal shards (``shard_range``): shards are padded to the largest one
    for the collective and the padding is dropped on ``dst``.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dets, counts
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    T, b_loc = counts.shape
    sizes = shard_sizes(n_images, world) if n_images is not None else [b_loc] * world
    b_max = max(sizes)
    if b_loc != sizes[rank]:
        raise ValueError(f"rank {rank} holds {b_loc} images, shard_range says {sizes[rank]}")
    if b_loc < b_max:  # pad the tail shard; masked by counts == 0
        pad_d = dets.new_zeros((T, b_max - b_loc) + tuple(dets.shape[2:]))
        pad_c = counts.new_zeros((T, b_max - b_loc))
        dets, counts = torch.cat((dets, pad_d), 1), torch.cat((counts, pad_c), 1)
    dets, counts = dets.contiguous(), counts.contiguous()
    if rank == dst:
        d_list = [torch.empty_like(dets) for _ in range(world)]
        c_list = [torch.empty_like(counts) for _ in range(world)]
    else:
        d_list = c_list = None
    dist.gather(dets, d_list, dst=dst, group=group)
    dist.gather(counts, c_list, dst=dst, group=group)
    if rank != dst:
        return None, None
    return (torch.cat([d[:, :n] for d, n in zip(d_list, sizes)], 1),
            torch.cat([c[:, :n] for c, n in zip(c_list, sizes)], 1))


class DetectionGatherer:
    """Persistent buffers for the per-batch gather: the NMS kernel writes ``dets`` and ``counts`` straight
    into one packed buffer per rank, so a batch needs ONE collective (``dist.gather`` of
    ``T*B*(max_det*6+1)`` 32-bit words) and no per-step allocation.  ``launch()`` is asynchronous; the
    gathered tensors on ``dst`` are valid after ``wait()``.  Equal shards only (pad the batch otherwise,
    or use ``gather_detections``)."""

    def __init__(self, T: int, b_loc: int, max_det: int, device, dst: int = 0, group=None):
        self.T, self.b_loc, self.max_det, self.dst, self.group = T, b_loc, max_det, dst, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        n_d, n_c = T * b_loc * max_det * 6, T * b_loc
        self.local = torch.empty(n_d + n_c, dtype=torch.float32, device=device)
        self.dets = self.local[:n_d].view(T, b_loc, max_det, 6)
        self.counts = self.local[n_d:].view(torch.int32).view(T, b_loc)
        self.all = None
        if self.rank == dst and self.world > 1:
            self.all = torch.empty((self.world, n_d + n_c), dtype=torch.float32, device=device)
            self._views = [self.all[r] for r in range(self.world)]
        self._n_d = n_d
        self._work = None

    @property
    def out(self):
        return self.dets, self.counts

    def launch(self):
        if self.world > 1:
            self._work = dist.gather(self.local, self._views if self.rank == self.dst else None, dst=self.dst,
                                     group=self.group, async_op=True)
        return self._work

    def wait(self):
        if self._work is not None:
            self._work.wait()
            self._work = None

    def result(self) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        """``(dets[T, world*B_loc, max_det, 6], counts[T, world*B_loc])`` on ``dst`` in global image order."""
        self.wait()
        if self.world == 1:
            return self.dets, self.counts
        if self.rank != self.dst:
            return None, None
        d = self.all[:, : self._n_d].view(self.world, self.T, self.b_loc, self.max_det, 6)
        c = self.all[:, self._n_d :].view(torch.int32).view(self.world, self.T, self.b_loc)
        return (d.permute(1, 0, 2, 3, 4).reshape(self.T, self.world * self.b_loc, self.max_det, 6),
                c.permute(1, 0, 2).reshape(self.T, self.world * self.b_loc))


# ------------------------------------------------------------------------------------------------ per-batch delivery to rank 0
class GatherDelivery:
    """Per-batch delivery of the padded detections to ``dst`` by one asynchronous ``dist.gather`` per batch
    (``DetectionGatherer``), two alternating buffers.  Works on every backend (NCCL on GPUs, gloo in the CPU tests)."""

    kind = "one asynchronous gather per batch (torch.distributed)"
    in_graph = False

    def __init__(self, T: int, b_loc: int, max_det: int, device, dst: int = 0, group=None):
        self.g = [DetectionGatherer(T, b_loc, max_det, device, dst=dst, group=group) for _ in range(2)]
        self.outs = [g.out for g in self.g]

    def before_write(self, slot: int) -> None:
        """Call before a kernel overwrites ``outs[slot]``: the previous batch in that slot must have left."""
        self.g[slot].wait()

    def after_write(self, slot: int) -> None:
        """Call after the kernel that filled ``outs[slot]`` was enqueued: starts the delivery of that batch."""
        self.g[slot].launch()

    def drain(self) -> None:
        for g in self.g:
            g.wait()

    def result(self, slot: int):
        return self.g[slot].result()

    def nms_deliver_args(self, slot: int):
        return None

    def collect(self, slot: int) -> None:
        pass


class PeerDelivery:
    """Per-batch delivery WITHOUT a collective and without a single extra launch on the writers: every rank's NMS kernel
    writes its padded detections straight into rank ``dst``'s memory over NVLink -- ``outs[slot]`` on rank r are views of
    rank dst's symmetric-memory buffer (``torch.distributed._symmetric_memory``: peer-mapped device memory), region
    (slot, r) -- and runs the hand-shake itself (``cerb_nms_deliver``, include/cerb_post.h):

      * its last CTA, after every CTA has fenced its remote stores, releases the slot's batch count into dst's flag word;
      * before its first store it checks that dst has acknowledged the batch it wrote into this slot two batches ago.

    Rank ``dst`` runs one 32-thread kernel per batch (``collect(slot)`` -> ``cerb_deliver_collect``): thread r waits for
    rank r's flag and stores the acknowledgement into rank r's memory.  All of it is stream-ordered kernels, so the
    pipeline captures it inside its CUDA graphs: per step the host replays one graph, on every rank.  No NCCL kernel takes
    SMs from the next batch's decode, no stream waits for a gather, nothing scales with the number of ranks on the host.

    Control words (uint32, after the data regions of the symmetric buffer; every rank allocates the same layout):
    ``flags[slot][r]`` at ``ctrl + 16*slot + r`` (used on dst), ``ack[slot]`` at ``ctrl + 32 + slot`` (used on writers).
    """

    kind = ("NMS kernel stores straight into rank 0's peer-mapped symmetric memory over NVLink and signals completion itself "
            "(flag / acknowledge words; no extra launch on the writers, one 32-thread collect kernel per batch on rank 0, "
            "all inside the step's CUDA graph)")
    in_graph = True
    MAX_WORLD = 16

    def __init__(self, T: int, b_loc: int, max_det: int, device, dst: int = 0, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        group = group if group is not None else dist.group.WORLD
        self.world, self.rank, self.dst = dist.get_world_size(group), dist.get_rank(group), dst
        if self.world > self.MAX_WORLD:
            raise ValueError(f"PeerDelivery supports at most {self.MAX_WORLD} ranks")
        self.T, self.b_loc, self.max_det = T, b_loc, max_det
        self.n_d, self.n_c = T * b_loc * max_det * 6, T * b_loc
        self.n = (self.n_d + self.n_c + 63) // 64 * 64  # words per (slot, rank) region, 256-byte aligned
        self.ctrl = 2 * self.world * self.n             # first control word
        try:
            symm_mem.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass  # newer builds enable it inside rendezvous()
        self.local = symm_mem.empty(self.ctrl + 64, dtype=torch.float32, device=device)
        self.local[self.ctrl :].zero_()
        self.hdl = symm_mem.rendezvous(self.local, group)
        self.outs = []
        for slot in range(2):
            off = (slot * self.world + self.rank) * self.n
            if self.rank == dst:
                region = self.local[off : off + self.n]
            else:
                region = self.hdl.get_buffer(dst, (self.n,), torch.float32, off)  # rank dst's memory, mapped here
            self.outs.append((region[: self.n_d].view(T, b_loc, max_det, 6),
                              region[self.n_d : self.n_d + self.n_c].view(torch.int32).view(T, b_loc)))
        # local protocol state: seq[2], done[2] (writers), collected[2] (dst)
        self.state = torch.zeros(8, dtype=torch.int32, device=device)
        base = self.local.data_ptr() + 4 * self.ctrl
        self._deliver = [None, None]
        self._collect = [None, None]
        self._peers = []  # keep the mapped views alive
        for slot in range(2):
            if self.rank != dst:
                flag = self.hdl.get_buffer(dst, (1,), torch.float32, self.ctrl + 16 * slot + self.rank)
                self._peers.append(flag)
                self._deliver[slot] = (flag.data_ptr(), base + 4 * (32 + slot), self.state.data_ptr() + 4 * slot,
                                       self.state.data_ptr() + 4 * (2 + slot))
            else:
                acks = []
                for r in range(self.world):
                    if r == dst:
                        acks.append(None)
                    else:
                        a = self.hdl.get_buffer(r, (1,), torch.float32, self.ctrl + 32 + slot)
                        self._peers.append(a)
                        acks.append(a.data_ptr())
                self._collect[slot] = (base + 4 * 16 * slot, acks, self.state.data_ptr() + 4 * (4 + slot))
        torch.cuda.synchronize(device)
        self.hdl.barrier(channel=7)

    def nms_deliver_args(self, slot: int):
        """``deliver=`` argument of ``ops.nms_batched`` for the NMS launch that fills ``outs[slot]`` (None on dst: its own
        rows are local and ordered by its stream)."""
        return self._deliver[slot]

    def collect(self, slot: int) -> None:
        """dst: enqueue (current stream) the kernel that waits for every rank's batch in ``slot`` and acknowledges it.
        A no-op on the other ranks."""
        if self.rank != self.dst:
            return
        from . import _lib

        flags, acks, collected = self._collect[slot]
        dev = self.local.device
        with torch.cuda.device(dev):
            _lib.check(_lib.load().cerb_deliver_collect(flags, _lib.ptr_array(acks), collected, self.world, self.dst,
                                                        torch.cuda.current_stream(dev).cuda_stream))

    # the hand-shake lives in the kernels: nothing to do around a write
    def before_write(self, slot: int) -> None:
        pass

    def after_write(self, slot: int) -> None:
        pass

    def drain(self) -> None:
        pass

    def result(self, slot: int):
        """``(dets[T, world*B_loc, max_det, 6], counts[T, world*B_loc])`` on ``dst`` (after the slot's ``collect`` has run
        and the device was synchronised), ``(None, None)`` elsewhere."""
        if self.rank != self.dst:
            return None, None
        reg = self.local[slot * self.world * self.n : (slot + 1) * self.world * self.n].view(self.world, self.n)
        d = reg[:, : self.n_d].reshape(self.world, self.T, self.b_loc, self.max_det, 6)
        c = reg[:, self.n_d : self.n_d + self.n_c].contiguous().view(torch.int32).view(self.world, self.T, self.b_loc)
        return (d.permute(1, 0, 2, 3, 4).reshape(self.T, self.world * self.b_loc, self.max_det, 6),
                c.permute(1, 0, 2).reshape(self.T, self.world * self.b_loc))


def make_delivery(T: int, b_loc: int, max_det: int, device, dst: int = 0, group=None, prefer_peer: bool = True):
    """The per-batch delivery to rank ``dst``: peer-mapped symmetric memory on CUDA when every rank can set it up
    (all ranks agree, so nobody is left in a collective alone), else the gather."""
    import os

    want_peer = prefer_peer and torch.device(device).type == "cuda" and os.environ.get("CERB_DELIVERY", "peer") != "gather"
    if want_peer:
        ok, deliv = 1, None
        try:
            deliv = PeerDelivery(T, b_loc, max_det, device, dst=dst, group=group)
        except Exception as exc:  # pragma: no cover - depends on the box / torch build
            import sys

            print(f"[cerberusdet_b200.shard] symmetric memory unavailable ({type(exc).__name__}: {exc}); using the gather", file=sys.stderr)
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 1:
            return deliv
    return GatherDelivery(T, b_loc, max_det, device, dst=dst, group=group)

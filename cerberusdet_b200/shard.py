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

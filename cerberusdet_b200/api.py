"""Host-buffer entry point of the path: raw head tensors in (pinned) host memory in,
padded detections in pinned host memory out.  This is what ``bench.py``'s ``e2e``
number times: H2D of the step's inputs, decode + NMS kernels, D2H of the result."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch

from . import ops


class HostPostProcessor:
    """Keeps device staging buffers and pinned output buffers between calls."""

    def __init__(self, device):
        self.device = torch.device(device)
        self._stage: Dict[tuple, torch.Tensor] = {}
        self._out: Dict[tuple, torch.Tensor] = {}

    def _staging(self, key, like):
        buf = self._stage.get(key)
        if buf is None or buf.shape != like.shape or buf.dtype != like.dtype:
            buf = torch.empty(like.shape, dtype=like.dtype, device=self.device)
            self._stage[key] = buf
        return buf

    def _pinned(self, key, like):
        buf = self._out.get(key)
        if buf is None or buf.shape != like.shape or buf.dtype != like.dtype:
            buf = torch.empty(like.shape, dtype=like.dtype, pin_memory=True)
            self._out[key] = buf
        return buf

    def __call__(self, heads_host: Sequence[Sequence[torch.Tensor]], strides, **nms_kw) -> Tuple[torch.Tensor, torch.Tensor]:
        with torch.cuda.device(self.device):
            dev_heads: List[List[torch.Tensor]] = []
            for t, lv in enumerate(heads_host):
                row = []
                for l, x in enumerate(lv):
                    buf = self._staging((t, l), x)
                    buf.copy_(x, non_blocking=True)
                    row.append(buf)
                dev_heads.append(row)
            ys = ops.decode_heads(dev_heads, strides)
            dets, counts = ops.nms_batched(ys, **nms_kw)
            h_dets = self._pinned("dets", dets)
            h_counts = self._pinned("counts", counts)
            h_dets.copy_(dets, non_blocking=True)
            h_counts.copy_(counts, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return h_dets, h_counts


_cache: Dict[str, HostPostProcessor] = {}


def postprocess_host(heads_host, strides, device="cuda:0", **nms_kw):
    """Decode + per-task NMS for host-resident raw head tensors ``heads_host[t][l]``.
    Returns pinned host tensors ``dets[T, B, max_det, 6]`` and ``counts[T, B]``."""
    key = str(torch.device(device))
    pp = _cache.get(key)
    if pp is None:
        pp = _cache[key] = HostPostProcessor(device)
    return pp(heads_host, strides, **nms_kw)

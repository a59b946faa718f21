"""Host restatement of the per-image statistics step that follows NMS in the reference validation loop
(SURVEY section 8f row 2): ``process_batch`` (cerberusdet/val.py:32-54) -- which detections count as correct at each
of the IoU thresholds ``iouv``.

Semantics, per threshold t: the pairs (label l, detection d) with ``iou >= t`` and equal class are sorted by IoU
descending; every detection keeps its first pair (its best label), the survivors -- now ordered by detection index --
are de-duplicated per label keeping the first, i.e. the lowest detection index (val.py:48-51).  The reference sorts with
numpy's unstable default, so exactly equal IoUs have no defined order there; the canonical rule here (and in the CUDA
kernel) is "IoU descending, then label index ascending".
"""
from __future__ import annotations

import torch

from .cross_task import pairwise_iou


def match_predictions(detections: torch.Tensor, labels: torch.Tensor, iouv: torch.Tensor) -> torch.Tensor:
    """``detections [N, 6]`` (x1, y1, x2, y2, conf, cls), ``labels [M, 5]`` (cls, x1, y1, x2, y2), ``iouv [K]``
    -> ``correct [N, K]`` bool (val.py:32-54)."""
    n, k = detections.shape[0], iouv.shape[0]
    correct = torch.zeros((n, k), dtype=torch.bool)
    if n == 0 or labels.shape[0] == 0:
        return correct
    det, lab = detections.detach().cpu().float(), labels.detach().cpu().float()
    iou = pairwise_iou(lab[:, 1:], det[:, :4])  # [M, N]
    same = lab[:, 0:1] == det[:, 5]
    masked = torch.where(same, iou, torch.full_like(iou, -1.0))
    best_iou, best_lab = masked.max(0)  # per detection; torch.max returns the first (lowest label) maximum
    iouv = iouv.detach().cpu().float()
    for i in range(k):
        valid = (best_iou >= iouv[i]) & (best_iou >= 0)
        taken = set()
        for d in range(n):  # ascending detection index: the first detection claiming a label keeps it
            if bool(valid[d]):
                l = int(best_lab[d])
                if l not in taken:
                    taken.add(l)
                    correct[d, i] = True
    return correct


def make_process_batch(reference_process_batch):
    """Drop-in for the reference's ``process_batch(detections, labels, iouv)`` (cerberusdet/val.py:32-54), bound onto
    ``cerberusdet.val`` by ``patch.install(val=True)``: the validation loop (val.py:321-357) calls it once per image.
    CUDA tensors go to ``cerb_val_match`` (ONE launch and no host round trip per image, where the reference makes
    ~5 launches and a ``.cpu().numpy()`` sync per IoU threshold -- 10 of them); anything else (CPU tensors, more than
    1024 labels or detections in an image) runs the reference's own function."""
    from . import ops

    cache = {}

    def process_batch(detections, labels, iouv):
        n, m = int(detections.shape[0]), int(labels.shape[0])
        on_path = (detections.is_cuda and labels.is_cuda and detections.dim() == 2 and detections.shape[1] == 6
                   and labels.dim() == 2 and labels.shape[1] == 5 and 0 < n <= 30000 and 0 < m <= 1024)
        if not on_path:
            return reference_process_batch(detections, labels, iouv)
        key = (iouv.data_ptr(), int(iouv.shape[0]), getattr(iouv, "_version", 0))
        host = cache.get(key)
        if host is None:  # the thresholds are one tensor for the whole run (val.py:206): read them once
            cache.clear()
            host = cache[key] = [float(v) for v in iouv.detach().cpu().float().tolist()]
        counts = torch.full((1,), n, dtype=torch.int32, device=detections.device)
        correct = ops.match_batch(detections.float().unsqueeze(0), counts, labels, [0, m], iouv, iouv_host=host)
        return correct[0]

    process_batch._cerb_reference = reference_process_batch
    return process_batch


def batch_statistics(dets: torch.Tensor, counts: torch.Tensor, labels: torch.Tensor, label_offsets, iouv: torch.Tensor):
    """The whole statistics step of one validation batch and one task in a single launch: ``dets [B, max_det, 6]`` /
    ``counts [B]`` straight from ``ops.nms_batched`` (already in native space, see ``ops.cross_task_merge``'s ``scale`` or
    the reference's ``scale_boxes``), ``labels [sum M_b, 5]`` native-space rows with ``label_offsets [B+1]``.
    Returns ``(correct [sum n_b, K] bool, conf [sum n_b], pcls [sum n_b])`` concatenated over the images in order -- the
    first three columns the reference appends to ``stats`` per image (val.py:357) -- as device tensors."""
    from . import ops

    correct = ops.match_batch(dets, counts, labels, label_offsets, iouv)  # [B, max_det, K]
    n = counts.to(torch.int64)
    keep = torch.arange(dets.shape[1], device=dets.device)[None, :] < n[:, None]  # rows below each image's count
    return correct[keep], dets[..., 4][keep], dets[..., 5][keep]

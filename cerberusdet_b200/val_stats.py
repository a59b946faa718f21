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

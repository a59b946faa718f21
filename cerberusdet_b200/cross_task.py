"""Host-side tail of ``CerberusDetInference.predict`` after the per-task NMS
(reference cerberusdet_inference.py:72-83,140-184): map local class ids to global ones,
suppress overlapping boxes of *different* tasks, rescale to the original image.

This is SURVEY section 8(f) row 1 ("next"): it stays host code for now, restated here so the
inference drop-in does not need the reference package at run time.  It mirrors
``nms_between_tasks`` (utils/general.py:484-554), ``scale_boxes`` / ``clip_boxes``
(utils/general.py:313-357) and ``box_iou`` (utils/metrics.py:415-433) operation for operation.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch


def pairwise_iou(a: torch.Tensor, b: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """utils/metrics.py:415-433: inter / (area_a + area_b - inter + eps) for xyxy boxes."""
    a1, a2 = a[:, None, :2], a[:, None, 2:4]
    b1, b2 = b[None, :, :2], b[None, :, 2:4]
    inter = (torch.min(a2, b2) - torch.max(a1, b1)).clamp(0).prod(2)
    return inter / ((a2 - a1).prod(2) + (b2 - b1).prod(2) - inter + eps)


def combine_tasks(per_task: Dict[str, torch.Tensor], id_maps: Dict[str, Dict[int, int]]) -> torch.Tensor:
    """cerberusdet_inference.py:72-83 for one image: concatenate the tasks' rows (CPU, fp32) with
    class ids mapped local -> global, in task order."""
    rows = [torch.zeros((0, 6))]
    for task, det in per_task.items():
        det = det.detach().cpu().clone()
        if det.shape[0]:
            lut = id_maps[task]
            det[:, 5] = torch.tensor([float(lut[int(c)]) for c in det[:, 5].tolist()])
            rows.append(det)
    return torch.cat(rows, 0)


def suppress_between_tasks(det: torch.Tensor, id_maps: Dict[str, Dict[int, int]], iou_thres: float) -> torch.Tensor:
    """utils/general.py:484-554.  Rows are first regrouped by task (in ``id_maps`` order); IoU is only
    taken between boxes of different tasks (block upper triangle).  Scanning rows top to bottom, an
    undeleted row with overlaps > thr keeps the arg-max score among {overlapping columns, itself} and
    deletes the others (already deleted columns still take part).  If every row would go, nothing does."""
    n = det.shape[0]
    groups: List[List[int]] = []
    for ids in id_maps.values():
        wanted = set(ids.values())
        groups.append([i for i in range(n) if int(det[i, 5]) in wanted])
    order = [i for g in groups for i in g]
    det = det[order]  # rows whose class belongs to no task are dropped, as in the reference
    m = det.shape[0]
    iou = torch.zeros((m, m))
    starts = [0]
    for g in groups:
        starts.append(starts[-1] + len(g))
    for i in range(len(groups)):
        if not groups[i]:
            continue
        for j in range(i + 1, len(groups)):
            if not groups[j]:
                continue
            iou[starts[i]:starts[i + 1], starts[j]:starts[j + 1]] = pairwise_iou(
                det[starts[i]:starts[i + 1], :4], det[starts[j]:starts[j + 1], :4])
    if not bool((iou > iou_thres).any()):
        return det
    gone = set()
    for r in range(m):
        if r in gone:
            continue
        cols = (iou[r] > iou_thres).nonzero().flatten().tolist()
        if not cols:
            continue
        cand = cols + [r]
        best = int(torch.argmax(det[cand, 4]))
        gone.update(c for k, c in enumerate(cand) if k != best)
    if len(gone) == m:
        return det
    keep = [i for i in range(m) if i not in gone]
    return det[keep]


def rescale_boxes(net_hw: Sequence[int], boxes: torch.Tensor, orig_hw: Sequence[int]) -> torch.Tensor:
    """utils/general.py:313-357 (ratio_pad=None): undo the letterbox, clip to the original image.
    In place, like the reference."""
    gain = min(net_hw[0] / orig_hw[0], net_hw[1] / orig_hw[1])
    pad_x = (net_hw[1] - orig_hw[1] * gain) / 2
    pad_y = (net_hw[0] - orig_hw[0] * gain) / 2
    boxes[..., [0, 2]] -= pad_x
    boxes[..., [1, 3]] -= pad_y
    boxes[..., :4] /= gain
    boxes[..., 0].clamp_(0, orig_hw[1])
    boxes[..., 1].clamp_(0, orig_hw[0])
    boxes[..., 2].clamp_(0, orig_hw[1])
    boxes[..., 3].clamp_(0, orig_hw[0])
    return boxes


def category_maps(names: Dict[str, List[str]]) -> Tuple[Dict[str, Dict[int, int]], List[str]]:
    """cerberusdet_inference.py:56-70: per-task local -> global class id maps, tasks laid end to end."""
    maps: Dict[str, Dict[int, int]] = {}
    all_names: List[str] = []
    base = 0
    for task, cats in names.items():
        maps[task] = {i: base + i for i in range(len(cats))}
        base += len(cats)
        all_names.extend(cats)
    return maps, all_names

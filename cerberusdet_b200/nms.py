"""Drop-in for the reference ``non_max_suppression`` (cerberusdet/utils/general.py:360-481)."""
from __future__ import annotations

from typing import List

import torch

from .ops import nms_batched


def non_max_suppression(
    prediction,
    conf_thres=0.25,
    iou_thres=0.45,
    classes=None,
    agnostic=False,
    multi_label=False,
    labels=(),
    max_det=300,
    nm=0,
) -> List[torch.Tensor]:
    """Same signature, argument meaning, asserts and return structure as the reference:
    a list of length B of fp32 ``[n_i, 6]`` tensors ``(x1, y1, x2, y2, conf, cls)`` on the
    input device, score-descending.  One kernel launch and one host sync (the counts)
    per call instead of ~25 launches and ~4 syncs per image.

    Differences, all deliberate: no wall-clock time limit (general.py:417,477-479 silently
    drops images); equal scores are ordered by (anchor, class) where the reference's
    unstable argsort leaves them unspecified; ``labels`` (autolabelling) and ``nm`` (masks)
    are not on this path and must be empty / 0.
    """
    if isinstance(prediction, (list, tuple)):  # (inference_out, loss_out), general.py:401-402
        prediction = prediction[0]
    if labels is not None and len(labels):
        raise NotImplementedError("apriori `labels` are not supported by the B200 path")
    if nm:
        raise NotImplementedError("mask coefficients (nm > 0) are not supported by the B200 path")
    dets, counts = nms_batched([prediction], conf_thres, iou_thres, classes, agnostic, multi_label, max_det)
    n = counts[0].tolist()  # the single device->host sync of the call
    return [dets[0, i, : n[i]] for i in range(len(n))]

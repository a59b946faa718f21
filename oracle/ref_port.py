"""CPU oracle for the CerberusDet post-head path (decode + per-task NMS).

TEST INFRASTRUCTURE ONLY.  Nothing under ``cerberusdet_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker or
the timed CPU baseline -- never as the product path.

It is a restatement of the reference algorithm, stage by stage, on ``torch`` CPU
tensors (the reference itself is pure torch, so the same ATen kernels do the
arithmetic).  Every function cites the reference lines it follows
(paths relative to the reference root).

Parity status: PINNED.  ``oracle/gen_golden.py`` runs the unmodified reference in
the build container and stores its outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this module against them bit for bit.

One deliberate difference from the reference: the candidate sort
(``utils/general.py:459``) is an *unstable* ``argsort`` there, so the order of
equal scores is unspecified.  The oracle fixes the canonical rule
"score descending, then candidate index ascending" (``stable=True``).  On tie-free
inputs both give identical results, which the golden generator asserts.

The greedy suppression itself lives in a third-party dependency that is not in
the reference tree: ``torchvision.ops.nms`` (pinned ``torchvision==0.20.1`` in
``pyproject.toml:52``; call site ``utils/general.py:464``).  ``greedy_nms.c`` restates
its published CPU algorithm in plain C; ``nms_port(..., greedy="c")`` uses that,
``greedy="torchvision"`` calls the installed binary, and the tests require both to
agree.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

REG_MAX = 16  # models/yolo.py:75  (DFL bins per box side)
MAX_WH = 7680  # utils/general.py:415
MAX_NMS = 30000  # utils/general.py:416

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_LIB = os.path.join(_BUILD, "libcerb_oracle.so")
_lib = None


# --------------------------------------------------------------------------- C part
def build_c_oracle(force: bool = False) -> str:
    """Compile ``greedy_nms.c`` (gcc, -O2, no FMA contraction) into ``oracle/_build``."""
    src = os.path.join(_HERE, "greedy_nms.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        os.makedirs(_BUILD, exist_ok=True)
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-o", _LIB, src]
        subprocess.run(cmd, check=True)
    return _LIB


def _c():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build_c_oracle())
        lib.oracle_greedy_nms.restype = ctypes.c_long
        lib.oracle_greedy_nms.argtypes = [
            ctypes.c_void_p,  # const float* boxes [n,4]
            ctypes.c_long,  # n
            ctypes.c_double,  # iou threshold
            ctypes.c_void_p,  # long* keep [n]
        ]
        lib.oracle_greedy_nms_topk.restype = ctypes.c_long
        lib.oracle_greedy_nms_topk.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_double, ctypes.c_void_p, ctypes.c_long]
        _lib = lib
    return _lib


def greedy_nms_c(boxes: torch.Tensor, iou_thres: float, max_keep: int = 0) -> torch.Tensor:
    """Greedy suppression over boxes ALREADY in processing order (score-descending).  ``max_keep > 0`` stops after that
    many kept boxes (the same first ``max_keep`` indices as the full run: greedy suppression is causal)."""
    b = np.ascontiguousarray(boxes.detach().cpu().numpy(), dtype=np.float32)
    n = b.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    k = _c().oracle_greedy_nms_topk(b.ctypes.data, n, float(iou_thres), keep.ctypes.data, int(max_keep))
    return torch.from_numpy(keep[:k].copy())


# --------------------------------------------------------------------------- decode
def anchor_grid(level_hw: Sequence[Sequence[int]], strides: Sequence[float], dtype, device="cpu"):
    """``make_anchors`` (utils/tal.py:181-193) + the transpose at models/yolo.py:94.

    Anchor k of a level with width w sits at (k % w + 0.5, k // w + 0.5) in grid
    units; levels are concatenated in order.  Returns ``anchors[2, A]`` and
    ``strides[1, A]`` in ``dtype``.
    """
    pts, st = [], []
    for (h, w), s in zip(level_hw, strides):
        xs = torch.arange(w, dtype=dtype, device=device) + 0.5
        ys = torch.arange(h, dtype=dtype, device=device) + 0.5
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        pts.append(torch.stack((gx, gy), -1).reshape(-1, 2))
        st.append(torch.full((h * w, 1), float(s), dtype=dtype, device=device))
    return torch.cat(pts).transpose(0, 1), torch.cat(st).transpose(0, 1)


def dfl_expectation(box_logits: torch.Tensor) -> torch.Tensor:
    """``DFL.forward`` (models/yolo.py:57-59): softmax over the 16 bins of each side,
    then a frozen 1x1 conv with weights 0..15 (models/yolo.py:52-54).

    ``box_logits`` is ``[B, 64, A]`` with channel = side*16 + bin.  The conv is kept
    as a conv (not a hand-written dot product) so half inputs round exactly where
    the reference's do.
    """
    b, _, a = box_logits.shape
    probs = box_logits.view(b, 4, REG_MAX, a).transpose(2, 1).softmax(1)  # [b,16,4,a]
    w = torch.arange(REG_MAX, dtype=torch.float).view(1, REG_MAX, 1, 1).to(box_logits)
    return F.conv2d(probs, w).view(b, 4, a)


def ltrb_to_xywh(dist: torch.Tensor, anchors: torch.Tensor) -> torch.Tensor:
    """``dist2bbox(xywh=True, dim=1)`` (utils/tal.py:196-205), operation order kept:
    corners first, then centre = (x1y1 + x2y2) / 2 and size = x2y2 - x1y1."""
    lt, rb = dist[:, :2], dist[:, 2:]
    p1 = anchors - lt
    p2 = anchors + rb
    return torch.cat(((p1 + p2) / 2, p2 - p1), 1)


def decode_port(levels: Sequence[torch.Tensor], nc: int, strides: Sequence[float]) -> torch.Tensor:
    """Eval branch of ``Detect.forward`` after the conv towers (models/yolo.py:93-99).

    ``levels[l]`` is the raw head tensor ``[B, 64+nc, H_l, W_l]``.  Returns
    ``y[B, 4+nc, A]`` in the input dtype: (cx, cy, w, h) in pixels, then sigmoid scores.
    """
    bsz = levels[0].shape[0]
    no = 4 * REG_MAX + nc
    dtype = levels[0].dtype
    anchors, stride_row = anchor_grid([t.shape[2:] for t in levels], strides, dtype, levels[0].device)
    flat = torch.cat([t.reshape(bsz, no, -1) for t in levels], 2)  # models/yolo.py:97
    box_logits, cls_logits = flat[:, : 4 * REG_MAX], flat[:, 4 * REG_MAX :]
    dbox = ltrb_to_xywh(dfl_expectation(box_logits), anchors.unsqueeze(0)) * stride_row  # yolo.py:98
    return torch.cat((dbox, cls_logits.sigmoid()), 1)  # yolo.py:99


# --------------------------------------------------------------------------- NMS
def centre_to_corners(b: torch.Tensor) -> torch.Tensor:
    """``xywh2xyxy`` (utils/general.py:272-288); arithmetic stays in the input dtype."""
    half_w = b[..., 2] / 2
    half_h = b[..., 3] / 2
    return torch.stack((b[..., 0] - half_w, b[..., 1] - half_h, b[..., 0] + half_w, b[..., 1] + half_h), -1)


def candidates_port(
    img: torch.Tensor,
    conf_thres: float,
    classes: Optional[Sequence[int]],
    multi_label: bool,
) -> torch.Tensor:
    """Stages N1-N3 for one image ``img[4+nc, A]`` -> fp32 rows (x1,y1,x2,y2,conf,cls).

    utils/general.py:411 (per-anchor prefilter, compare in tensor dtype), :427
    (rows in ascending anchor order), :443 (corners), :444-449 (multi-label
    expansion in anchor-major, class-minor order, or best class with the lowest
    index winning ties), :452-453 (class filter).
    """
    nc = img.shape[0] - 4
    rows = img.transpose(0, 1)  # [A, 4+nc]
    rows = rows[rows[:, 4:].amax(1) > conf_thres]
    if rows.shape[0] == 0:
        return torch.zeros((0, 6))
    box = centre_to_corners(rows[:, :4])
    scores = rows[:, 4:]
    if multi_label and nc > 1:  # utils/general.py:419
        r, c = (scores > conf_thres).nonzero(as_tuple=True)
        det = torch.cat((box[r].float(), scores[r, c].float()[:, None], c.float()[:, None]), 1)
    else:
        conf, c = scores.max(1)
        det = torch.cat((box.float(), conf.float()[:, None], c.float()[:, None]), 1)
        det = det[conf > conf_thres]
    if classes is not None:
        wanted = torch.tensor(list(classes), dtype=torch.float32)
        det = det[(det[:, 5:6] == wanted).any(1)]
    return det


def nms_port(
    prediction: torch.Tensor,
    conf_thres: float = 0.25,
    iou_thres: float = 0.45,
    classes: Optional[Sequence[int]] = None,
    agnostic: bool = False,
    multi_label: bool = False,
    max_det: int = 300,
    max_nms: int = MAX_NMS,
    max_wh: float = MAX_WH,
    greedy: str = "torchvision",
    return_candidates: bool = False,
) -> List[torch.Tensor]:
    """``non_max_suppression`` (utils/general.py:360-481) with ``labels=()``, ``nm=0``.

    No wall-clock time limit (utils/general.py:417,477-479 is a result hazard, not
    part of the algorithm) and a stable candidate sort (see module docstring).
    """
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}"  # :399
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}"  # :400
    if isinstance(prediction, (list, tuple)):  # :401-402
        prediction = prediction[0]
    out: List[torch.Tensor] = []
    cands = []
    for img in prediction:
        det = candidates_port(img, conf_thres, classes, multi_label)
        if det.shape[0]:
            order = det[:, 4].argsort(descending=True, stable=True)[:max_nms]  # :459 (canonical ties)
            det = det[order]
            shifted = det[:, :4] + det[:, 5:6] * (0 if agnostic else max_wh)  # :462-463
            if greedy == "torchvision":
                import torchvision

                keep = torchvision.ops.nms(shifted, det[:, 4], iou_thres)  # :464
            else:
                keep = greedy_nms_c(shifted, iou_thres, max_keep=max_det)  # early exit: same first max_det indices
            det_out = det[keep[:max_det]]  # :465,474
        else:
            det_out = det
        out.append(det_out)
        cands.append(det)
    return (out, cands) if return_candidates else out


def postprocess_port(task_levels, ncs, strides, **nms_kw):
    """Whole path for several task heads: decode each head, then NMS each head
    (the per-task loop of cerberusdet_inference.py:125-135)."""
    results = []
    for levels, nc in zip(task_levels, ncs):
        y = decode_port(levels, nc, strides)
        results.append(nms_port(y, **nms_kw))
    return results


# --------------------------------------------------------------------------- training-time sibling decode
def bbox_decode_port(anchor_points: torch.Tensor, pred_dist: torch.Tensor) -> torch.Tensor:
    """``Loss.bbox_decode`` with ``use_dfl`` (utils/loss.py:126-131) followed by ``dist2bbox(xywh=False)``
    (utils/tal.py:196-205): ``pred_dist [B, A, 64]`` (bins of a side contiguous) and ``anchor_points [A, 2]`` ->
    ``[B, A, 4]`` = (x1, y1, x2, y2) in grid units.  Differentiable (the test compares autograd gradients too)."""
    b, a, c = pred_dist.shape
    proj = torch.arange(REG_MAX, dtype=torch.float)
    dist = pred_dist.view(b, a, 4, c // 4).softmax(3).matmul(proj.type(pred_dist.dtype))
    lt, rb = torch.split(dist, 2, -1)
    return torch.cat((anchor_points - lt, anchor_points + rb), -1)


# --------------------------------------------------------------------------- head tail (the checker is ready, the kernel is round 2)
def head_tail_port(box_feats: Sequence[torch.Tensor], cls_feats: Sequence[torch.Tensor], box_w: Sequence[torch.Tensor],
                   box_b: Sequence[torch.Tensor], cls_w: Sequence[torch.Tensor], cls_b: Sequence[torch.Tensor],
                   strides: Sequence[float]) -> Tuple[torch.Tensor, list]:
    """The step right before the path plus the path: the LAST 1x1 convolutions of the two towers of every level
    (``cv2[i][-1]``: c2 -> 64 and ``cv3[i][-1]``: c3 -> nc, models/yolo.py:81-84), the channel concat (yolo.py:89-90) and
    the eval decode (yolo.py:93-99).

    ``box_feats[l]`` ``[B, c2, H_l, W_l]`` / ``cls_feats[l]`` ``[B, c3, H_l, W_l]`` are the inputs of those last
    convolutions; ``box_w[l]`` ``[64, c2, 1, 1]``, ``cls_w[l]`` ``[nc, c3, 1, 1]`` and the biases are their parameters.
    Returns ``(y [B, 4+nc, A], raw levels [B, 64+nc, H_l, W_l])`` like ``Detect.forward`` in eval mode."""
    raw = []
    for l in range(len(box_feats)):
        raw.append(torch.cat((F.conv2d(box_feats[l], box_w[l], box_b[l]), F.conv2d(cls_feats[l], cls_w[l], cls_b[l])), 1))
    nc = int(cls_w[0].shape[0])
    return decode_port(raw, nc, strides), raw

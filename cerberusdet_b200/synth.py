"""Synthetic head tensors and predictions for tests and ``bench.py``.

There are no datasets or checkpoints offline, so the workloads named in
``BASELINE.json`` are driven by seeded synthetic inputs.
Every image has its own seed, so any sharding of a batch over ranks reproduces
exactly the same global batch.

Regimes for raw head tensors ``[64+nc, H_l, W_l]`` (what the Detect conv towers emit,
reference ``models/yolo.py:89-90``):

* ``iid``     box logits N(0, 3^2), class logits N(-6, 2^2)  (configs 2-5)
* ``planted`` background class logits N(-8, 1) plus K objects per image whose
  centre anchors get a high class logit and DFL logits peaked at the true
  distances -- gives overlapping clusters like a trained detector's output.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

REG_MAX = 16
STRIDES = (8.0, 16.0, 32.0)  # P3/P4/P5, reference models/yolov8x_voc_obj365.yaml:42
TASK_NC = {"voc": 20, "objects365_animals": 19, "objects365_tableware": 12}  # reference data/*.yaml


def level_shapes(imgsz, strides: Sequence[float] = STRIDES) -> List[Tuple[int, int]]:
    h, w = (imgsz, imgsz) if isinstance(imgsz, int) else imgsz
    return [(int(h // s), int(w // s)) for s in strides]


def image_seed(cfg: int, image: int, task: int, level: int) -> int:
    return 1_000_003 * cfg + 10_007 * image + 101 * task + level


def _planted_level(gen, nc, h, w, stride, objects):
    """One level of one image in the ``planted`` regime."""
    x = torch.empty(4 * REG_MAX + nc, h, w)
    x[: 4 * REG_MAX].normal_(0.0, 1.0, generator=gen)
    x[4 * REG_MAX :].normal_(-8.0, 1.0, generator=gen)
    ys = (torch.arange(h, dtype=torch.float32) + 0.5) * stride
    xs = (torch.arange(w, dtype=torch.float32) + 0.5) * stride
    bins = torch.arange(REG_MAX, dtype=torch.float32)
    for (x1, y1, x2, y2, cls, lvl_stride) in objects:
        if lvl_stride != stride:
            continue
        cx, cy, bw, bh = (x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1
        my = ((ys - cy).abs() <= bh / 4).nonzero().flatten()
        mx = ((xs - cx).abs() <= bw / 4).nonzero().flatten()
        for iy in my.tolist():
            for ix in mx.tolist():
                x[4 * REG_MAX + cls, iy, ix] = 2.0 + float(torch.randn((), generator=gen))
                d = torch.tensor([xs[ix] - x1, ys[iy] - y1, x2 - xs[ix], y2 - ys[iy]]) / stride
                d = d.clamp(0, REG_MAX - 1.01)
                for side in range(4):
                    peak = 8.0 * (1 - (bins - d[side]).abs()).clamp_min(0)
                    x[side * REG_MAX : (side + 1) * REG_MAX, iy, ix] += peak
    return x


def synth_heads(
    batch_images: Sequence[int],
    ncs: Sequence[int],
    imgsz=640,
    dtype=torch.float16,
    regime: str = "iid",
    cfg: int = 3,
    strides: Sequence[float] = STRIDES,
    pin: bool = False,
) -> List[List[torch.Tensor]]:
    """Raw head tensors ``out[task][level] = [B, 64+nc, H_l, W_l]`` (CPU) for the
    global image indices ``batch_images``."""
    shapes = level_shapes(imgsz, strides)
    bsz = len(batch_images)
    out = []
    for t, nc in enumerate(ncs):
        no = 4 * REG_MAX + nc
        lv = []
        for (h, w) in shapes:
            buf = torch.empty((bsz, no, h, w), dtype=dtype)
            lv.append(buf.pin_memory() if pin else buf)
        out.append(lv)
    for bi, img in enumerate(batch_images):
        for t, nc in enumerate(ncs):
            objects = None
            if regime == "planted":
                g0 = torch.Generator().manual_seed(image_seed(cfg, img, t, 99))
                k = int(torch.randint(5, 41, (1,), generator=g0))
                H = shapes[0][0] * strides[0]
                W = shapes[0][1] * strides[0]
                objects = []
                for _ in range(k):
                    bw = float(torch.empty(1).uniform_(24, 0.6 * W, generator=g0))
                    bh = float(torch.empty(1).uniform_(24, 0.6 * H, generator=g0))
                    cx = float(torch.empty(1).uniform_(bw / 2, W - bw / 2, generator=g0))
                    cy = float(torch.empty(1).uniform_(bh / 2, H - bh / 2, generator=g0))
                    cls = int(torch.randint(0, nc, (1,), generator=g0))
                    size = max(bw, bh)
                    s = strides[0] if size < 96 else (strides[1] if size < 224 else strides[2])
                    objects.append((cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2, cls, s))
            for l, (h, w) in enumerate(shapes):
                gen = torch.Generator().manual_seed(image_seed(cfg, img, t, l))
                if regime == "iid":
                    x = torch.empty(4 * REG_MAX + nc, h, w)
                    x[: 4 * REG_MAX].normal_(0.0, 3.0, generator=gen)
                    x[4 * REG_MAX :].normal_(-6.0, 2.0, generator=gen)
                elif regime == "planted":
                    x = _planted_level(gen, nc, h, w, strides[l], objects)
                else:
                    raise ValueError(f"unknown regime {regime!r}")
                out[t][l][bi].copy_(x)
    return out


def synth_prediction(
    bsz: int,
    nc: int,
    anchors: int,
    seed: int,
    dtype=torch.float32,
    regime: str = "clusters",
    imgsz: float = 640.0,
    score_scale: float = 1.0,
) -> torch.Tensor:
    """A decoded prediction ``[B, 4+nc, A]`` (the input of ``non_max_suppression``).

    ``clusters``: boxes jittered around a few dozen centres per image so that NMS
    really suppresses; scores are sigmoid(N(-4, 2.5^2)) * score_scale.
    ``uniform``: independent boxes.
    """
    g = torch.Generator().manual_seed(seed)
    y = torch.empty(bsz, 4 + nc, anchors)
    if regime == "clusters":
        k = 24
        centres = torch.rand(bsz, 2, k, generator=g) * imgsz
        sizes = 20 + torch.rand(bsz, 2, k, generator=g) * 0.4 * imgsz
        pick = torch.randint(0, k, (bsz, anchors), generator=g)
        idx = pick[:, None, :].expand(bsz, 2, anchors)
        y[:, 0:2] = centres.gather(2, idx) + torch.randn(bsz, 2, anchors, generator=g) * 6
        y[:, 2:4] = sizes.gather(2, idx) * (1 + 0.15 * torch.randn(bsz, 2, anchors, generator=g)).clamp_min(0.05)
    elif regime == "uniform":
        y[:, 0:2] = torch.rand(bsz, 2, anchors, generator=g) * imgsz
        y[:, 2:4] = 4 + torch.rand(bsz, 2, anchors, generator=g) * 0.3 * imgsz
    else:
        raise ValueError(regime)
    y[:, 4:] = torch.sigmoid(torch.randn(bsz, nc, anchors, generator=g) * 2.5 - 4.0) * score_scale
    return y.to(dtype)

"""Anchor / stride tensors in the reference's cached layout (``Detect.anchors [2, A]``,
``Detect.strides [1, A]``; reference utils/tal.py:181-193 + models/yolo.py:94).  The CUDA
kernels compute anchors analytically; these tensors exist only because other reference code
reads the attributes."""
from __future__ import annotations

import torch


def make_anchor_tensors(level_hw, strides, dtype, device):
    pts, st = [], []
    for (h, w), s in zip(level_hw, strides):
        xs = torch.arange(int(w), dtype=dtype, device=device) + 0.5
        ys = torch.arange(int(h), dtype=dtype, device=device) + 0.5
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        pts.append(torch.stack((gx, gy), -1).reshape(-1, 2))
        st.append(torch.full((int(h) * int(w), 1), float(s), dtype=dtype, device=device))
    return torch.cat(pts).transpose(0, 1), torch.cat(st).transpose(0, 1)

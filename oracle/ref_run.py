"""Run the UNMODIFIED reference's own code for the post-head path on given raw head tensors (TEST INFRASTRUCTURE ONLY;
used by tests/ and by bench.py's reference legs -- ``--impl reference``, ``cpu_baseline`` and the same-GPU eager baseline).

Decode: the reference ``Detect`` class (models/yolo.py:64-100) with each level's conv towers replaced by channel slices, so
its eval branch (:93-100) runs verbatim on a given raw head tensor (SURVEY 8c "slice-oracle").  NMS: the reference
``non_max_suppression`` (utils/general.py:360-481) called ONE image per call (its wall-clock limit, :417/:477-479, can
otherwise drop images).  Works on CPU tensors and, on a GPU box, on CUDA tensors (torchvision's CUDA kernel then).
The reference tree is found by ``oracle/ref_import.py`` (/root/reference, or its copy oracle/_ref).
"""
from __future__ import annotations

import warnings

import torch
import torch.nn as nn

from .ref_import import load_reference, reference_available  # noqa: F401


class _Slice(nn.Module):
    def __init__(self, lo, hi):
        super().__init__()
        self.lo, self.hi = lo, hi

    def forward(self, x):
        return x[:, self.lo : self.hi]


class ReferenceRunner:
    def __init__(self, strides=(8.0, 16.0, 32.0)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            self.ref = load_reference()
        self.strides = strides
        self._heads = {}

    def head(self, nc, dtype, device):
        key = (nc, dtype, str(device))
        m = self._heads.get(key)
        if m is None:
            m = self.ref.yolo.Detect(nc=nc, ch=(16,) * len(self.strides))
            m.stride = torch.tensor(self.strides)
            for i in range(len(self.strides)):
                m.cv2[i] = _Slice(0, 64)
                m.cv3[i] = _Slice(64, 64 + nc)
            m = m.to(device).eval()
            if dtype == torch.float16:
                m.half()
            self._heads[key] = m
        return m

    def decode(self, levels, nc):
        """Reference Detect.forward eval branch on raw head tensors ``levels[l] [B, 64+nc, H_l, W_l]`` -> y."""
        m = self.head(nc, levels[0].dtype, levels[0].device)
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            y, _ = m([t.clone() for t in levels])
        return y

    def nms(self, pred, **kw):
        """Reference non_max_suppression, one image per call -> list of [n_i, 6] tensors."""
        return [self.ref.general.non_max_suppression(pred[i : i + 1], **kw)[0] for i in range(pred.shape[0])]

    def segment(self, levels, nc, **kw):
        """decode + NMS of one head on (a slice of) a batch."""
        return self.nms(self.decode(levels, nc), **kw)

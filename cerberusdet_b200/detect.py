"""Drop-in for the eval branch of the reference ``Detect.forward``
(cerberusdet/models/yolo.py:87-100).

The reference class must stay the class that checkpoints were pickled with
(utils/models_manager.py:296-308, models/experimental.py:112-122), so nothing here
subclasses it: ``detect_forward`` is bound onto the existing class by
``cerberusdet_b200.patch.install()``.  It takes over exactly one thing -- the decode of
CUDA tensors in eval mode; training mode and non-CUDA tensors run the reference's own
method untouched.
"""
from __future__ import annotations

import torch

from .ops import decode_heads

_RAW_FLAG = "_cerb_raw_heads"  # set by inference.raw_heads(): return the per-level tensors undecoded
_STRIDES_ATTR = "_cerb_strides"  # (key, python floats): the head's strides without a device sync per forward


def head_strides(head):
    """``head.stride`` as Python floats.  ``attempt_load(..., map_location=device)`` leaves that tensor on the GPU, where
    every ``float()`` is a blocking device->host copy behind the queued conv towers -- so the floats are read once and
    kept on the module, keyed by the tensor object and its version (a replaced or rewritten stride is re-read)."""
    st = head.stride
    key = (id(st), getattr(st, "_version", 0)) if torch.is_tensor(st) else None
    ent = getattr(head, _STRIDES_ATTR, None)
    if ent is None or ent[0] != key or key is None:
        vals = tuple(float(v) for v in (st.detach().cpu().tolist() if torch.is_tensor(st) else st))
        ent = (key, vals)
        object.__setattr__(head, _STRIDES_ATTR, ent)
    return ent[1]


def _level_cat(self, x):
    # the conv towers belong to the model, not to this path (yolo.py:89-90)
    for i in range(self.nl):
        x[i] = torch.cat((self.cv2[i](x[i]), self.cv3[i](x[i])), 1)
    return x


class SplitHeads:
    """Raw heads of one task with the channel concat of models/yolo.py:89-90 left out: ``box[l]`` = cv2[l]'s output
    ``[B, 64, H_l, W_l]``, ``cls[l]`` = cv3[l]'s ``[B, nc, H_l, W_l]``.  ``ops.decode_heads_split`` reads both."""

    __slots__ = ("box", "cls")

    def __init__(self, box, cls):
        self.box, self.cls = box, cls


def detect_forward(self, x):
    """``forward(self, x: list[Tensor]) -> (y, x)`` | ``y`` (export) | ``x`` (training).

    ``y`` is ``[B, 4+nc, A]`` in the input dtype, rows cx, cy, w, h (pixels) then sigmoid
    scores -- the reference layout.  ``x`` is mutated in place with the per-level
    concatenations exactly like the reference does (yolo.py:89-90).
    """
    reference_forward = type(self)._cerb_reference_forward
    if self.training or not x[0].is_cuda or x[0].dtype not in (torch.float16, torch.float32):
        return reference_forward(self, x)
    shape = x[0].shape  # BCHW before the convs: the reference's anchor-cache key (yolo.py:88,93)
    if getattr(self, _RAW_FLAG, False):
        # inference engine: nobody reads the concatenated raw tensors, so the towers' outputs are handed over as they
        # are and the decode kernel reads box and class channels from their own tensors (no torch.cat copy)
        return None, SplitHeads([self.cv2[i](x[i]) for i in range(self.nl)], [self.cv3[i](x[i]) for i in range(self.nl)])
    x = _level_cat(self, x)
    if self.dynamic or self.shape != shape:
        # keep the instance attributes other code reads (and pickles) populated (yolo.py:93-95)
        from .anchors import make_anchor_tensors

        self.anchors, self.strides = make_anchor_tensors([t.shape[2:] for t in x], self.stride, x[0].dtype, x[0].device)
        self.shape = shape
    strides = head_strides(self)
    y = decode_heads([[t if t.is_contiguous() else t.contiguous() for t in x]], strides)[0]
    return y if self.export else (y, x)

"""``install()`` -- make an existing CerberusDet checkout use the B200 path.

What gets rebound (and restored by ``uninstall()``):

* ``cerberusdet.models.yolo.Detect.forward``  -> ``detect.detect_forward`` (the class itself is kept:
  checkpoints pickle it by qualified name);
* the name ``non_max_suppression`` in ``cerberusdet.utils.general`` and in every module that imported it
  by name at import time (``val``, ``detect``, ``cerberusdet_inference`` -- reference val.py:17,
  detect.py:18, cerberusdet_inference.py:10);
* ``cerberusdet.cerberusdet_inference.CerberusDetInference`` -> ``inference.CerberusDetInference``;
* with ``train=True`` also ``cerberusdet.utils.loss.Loss.bbox_decode`` -> ``ops.bbox_decode`` (the training-time
  sibling of the decode, reference utils/loss.py:126-131; forward and backward kernels).

CUDA fp16/fp32 tensors go to the kernels; anything else (CPU tensors, training mode, masks/labels
arguments) is handed to the reference's own, saved implementation -- the patch never changes what a
non-CUDA run computes.
"""
from __future__ import annotations

import importlib
import sys
from typing import Dict, List, Tuple

_saved: List[Tuple[object, str, object]] = []
_NMS_IMPORTERS = ("cerberusdet.utils.general", "cerberusdet.val", "cerberusdet.detect", "cerberusdet.cerberusdet_inference")


def _set(obj, name, value):
    _saved.append((obj, name, getattr(obj, name)))
    setattr(obj, name, value)


def installed() -> bool:
    return bool(_saved)


def install(import_all: bool = False, train: bool = False) -> Dict[str, List[str]]:
    """Patch the reference modules that are importable.  ``import_all`` also imports ``val`` /
    ``detect`` / ``cerberusdet_inference`` (they pull in the whole data pipeline); by default only
    modules already imported, plus ``models.yolo`` and ``utils.general``, are touched.  ``train`` also rebinds
    ``Loss.bbox_decode`` (imports ``cerberusdet.utils.loss``)."""
    if _saved:
        return {"already": ["installed"]}
    import torch

    from . import detect as _detect
    from . import inference as _inference
    from . import nms as _nms

    done: Dict[str, List[str]] = {"patched": []}
    yolo = importlib.import_module("cerberusdet.models.yolo")
    general = importlib.import_module("cerberusdet.utils.general")

    det_cls = yolo.Detect
    if not hasattr(det_cls, "_cerb_reference_forward"):
        det_cls._cerb_reference_forward = det_cls.forward
    _set(det_cls, "forward", _detect.detect_forward)
    done["patched"].append("cerberusdet.models.yolo.Detect.forward")

    reference_nms = general.non_max_suppression

    def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                            multi_label=False, labels=(), max_det=300, nm=0):
        p = prediction[0] if isinstance(prediction, (list, tuple)) else prediction
        on_path = p.is_cuda and p.dtype in (torch.float16, torch.float32) and not nm and not (labels is not None and len(labels))
        if not on_path:
            return reference_nms(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, labels, max_det, nm)
        return _nms.non_max_suppression(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, labels, max_det, nm)

    non_max_suppression.__doc__ = _nms.non_max_suppression.__doc__
    non_max_suppression._cerb_reference = reference_nms
    for modname in _NMS_IMPORTERS:
        mod = sys.modules.get(modname)
        if mod is None and (import_all or modname == "cerberusdet.utils.general"):
            mod = importlib.import_module(modname)
        if mod is not None and hasattr(mod, "non_max_suppression"):
            _set(mod, "non_max_suppression", non_max_suppression)
            done["patched"].append(f"{modname}.non_max_suppression")

    inf = sys.modules.get("cerberusdet.cerberusdet_inference")
    if inf is None and import_all:
        inf = importlib.import_module("cerberusdet.cerberusdet_inference")
    if inf is not None:
        _set(inf, "CerberusDetInference", _inference.CerberusDetInference)
        done["patched"].append("cerberusdet.cerberusdet_inference.CerberusDetInference")
    if train:
        from . import ops as _ops

        loss_mod = importlib.import_module("cerberusdet.utils.loss")
        reference_bbox_decode = loss_mod.Loss.bbox_decode

        def bbox_decode(self, anchor_points, pred_dist):
            on_path = (self.use_dfl and pred_dist.is_cuda and pred_dist.dtype in (torch.float16, torch.float32)
                       and pred_dist.dim() == 3 and pred_dist.shape[-1] == 64 and not torch.is_autocast_enabled())
            if not on_path:  # CPU tensors, reg_max != 16, autocast (its softmax/matmul casts are the reference's business)
                return reference_bbox_decode(self, anchor_points, pred_dist)
            return _ops.bbox_decode(anchor_points, pred_dist)

        bbox_decode._cerb_reference = reference_bbox_decode
        _set(loss_mod.Loss, "bbox_decode", bbox_decode)
        done["patched"].append("cerberusdet.utils.loss.Loss.bbox_decode")
    return done


def uninstall() -> None:
    while _saved:
        obj, name, value = _saved.pop()
        setattr(obj, name, value)

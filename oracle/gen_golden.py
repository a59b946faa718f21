"""Generate ``tests/golden/*.npz`` by executing the UNMODIFIED reference.

Run in the build container only (it needs ``/root/reference``):

    python -m oracle.gen_golden

Decode vectors come from the reference ``Detect.forward`` (models/yolo.py:87-100)
with each level's conv towers replaced by channel slices, so the eval branch
(:93-100) runs verbatim on a given raw head tensor.  NMS vectors come from the
reference ``non_max_suppression`` (utils/general.py:360-481) called ONE image per
call (its wall-clock limit, :417/:477-479, can otherwise drop images).

The reference's candidate sort (:459) is an unstable argsort; for every NMS vector
the function is run twice -- as is, and with ``Tensor.argsort`` forced to
``stable=True`` -- and the stable result is stored.  ``plain_equal`` records whether
the as-is run gave the same rows (it must whenever the scores that matter are
tie-free; for fp32 vectors the script asserts that any difference is a permutation
of rows with exactly equal scores).
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from cerberusdet_b200.synth import STRIDES, synth_heads, synth_prediction  # noqa: E402
from oracle.ref_import import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class _Slice(nn.Module):
    def __init__(self, lo, hi):
        super().__init__()
        self.lo, self.hi = lo, hi

    def forward(self, x):
        return x[:, self.lo : self.hi]


def reference_decode(ref, levels, nc):
    """Run the reference Detect eval branch on raw head tensors (slice-oracle)."""
    m = ref.yolo.Detect(nc=nc, ch=(16, 16, 16))
    m.stride = torch.tensor(STRIDES)
    for i in range(3):
        m.cv2[i] = _Slice(0, 64)
        m.cv3[i] = _Slice(64, 64 + nc)
    m.eval()
    if levels[0].dtype == torch.float16:
        m.half()
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        y, _ = m([t.clone() for t in levels])
    return y


class _StableArgsort:
    """Force stable=True on Tensor.argsort while the reference function runs."""

    def __enter__(self):
        self.orig = torch.Tensor.argsort

        def stable(t, *a, **k):
            k["stable"] = True
            return self.orig(t, *a, **k)

        torch.Tensor.argsort = stable

    def __exit__(self, *exc):
        torch.Tensor.argsort = self.orig


def reference_nms(ref, pred, stable, **kw):
    outs = []
    for i in range(pred.shape[0]):
        if stable:
            with _StableArgsort():
                o = ref.general.non_max_suppression(pred[i : i + 1], **kw)[0]
        else:
            o = ref.general.non_max_suppression(pred[i : i + 1], **kw)[0]
        outs.append(o.clone())
    return outs


def _pack(rows):
    counts = np.array([r.shape[0] for r in rows], dtype=np.int64)
    cat = torch.cat([r.reshape(-1, 6) for r in rows], 0).numpy() if rows else np.zeros((0, 6), np.float32)
    return cat.astype(np.float32), counts


def _np(t):
    return t.numpy()


DECODE_CASES = [
    # name, nc, imgsz(h,w), B, dtype, regime, logit scale
    ("decode_f32_nc20_64x96", 20, (64, 96), 2, torch.float32, "iid", 1.0),
    ("decode_f16_nc19_128", 19, (128, 128), 3, torch.float16, "iid", 1.0),
    ("decode_f32_nc12_320_planted", 12, (320, 320), 1, torch.float32, "planted", 1.0),
    ("decode_f16_nc20_640", 20, (640, 640), 1, torch.float16, "iid", 1.0),
    ("decode_f16_nc3_96x160_wide", 3, (96, 160), 2, torch.float16, "iid", 6.0),
    ("decode_f32_nc1_64_wide", 1, (64, 64), 2, torch.float32, "iid", 10.0),
]

# name, dict(pred=...), nms kwargs
NMS_CASES = [
    ("nms_f32_best_2100", dict(bsz=2, nc=20, anchors=2100, seed=11, dtype=torch.float32, regime="clusters"),
     dict(conf_thres=0.25, iou_thres=0.45)),
    ("nms_f32_multi_2100", dict(bsz=2, nc=12, anchors=2100, seed=12, dtype=torch.float32, regime="clusters"),
     dict(conf_thres=0.001, iou_thres=0.6, multi_label=True)),
    ("nms_f32_multi_8400_trunc", dict(bsz=1, nc=20, anchors=8400, seed=13, dtype=torch.float32, regime="clusters"),
     dict(conf_thres=0.001, iou_thres=0.6, multi_label=True)),
    ("nms_f16_best_2100", dict(bsz=2, nc=19, anchors=2100, seed=14, dtype=torch.float16, regime="clusters"),
     dict(conf_thres=0.3, iou_thres=0.45)),
    ("nms_f16_multi_2100", dict(bsz=2, nc=20, anchors=2100, seed=15, dtype=torch.float16, regime="clusters"),
     dict(conf_thres=0.001, iou_thres=0.6, multi_label=True)),
    ("nms_f32_classes_agnostic", dict(bsz=2, nc=20, anchors=1000, seed=16, dtype=torch.float32, regime="clusters"),
     dict(conf_thres=0.1, iou_thres=0.5, classes=[0, 3, 7, 19], agnostic=True, max_det=50)),
    ("nms_f32_multi_classes_maxdet5", dict(bsz=2, nc=12, anchors=1000, seed=17, dtype=torch.float32, regime="uniform"),
     dict(conf_thres=0.05, iou_thres=0.3, classes=[1, 2, 11], multi_label=True, max_det=5)),
    ("nms_f32_uniform_maxdet1000", dict(bsz=1, nc=20, anchors=8400, seed=18, dtype=torch.float32, regime="uniform"),
     dict(conf_thres=0.02, iou_thres=0.45, max_det=1000)),
    ("nms_f16_saturated", dict(bsz=2, nc=5, anchors=600, seed=19, dtype=torch.float16, regime="clusters", score_scale=1.0),
     dict(conf_thres=0.25, iou_thres=0.45, multi_label=True)),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    torch.set_num_threads(os.cpu_count())
    manifest = {}

    for name, nc, imgsz, bsz, dtype, regime, scale in DECODE_CASES:
        heads = synth_heads(range(bsz), [nc], imgsz, torch.float32, regime, cfg=7)[0]
        heads = [(h * scale).to(dtype) for h in heads]
        y = reference_decode(ref, heads, nc)
        assert y.dtype == dtype and torch.isfinite(y.float()).all()
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            **{f"level{i}": _np(h) for i, h in enumerate(heads)},
            y=_np(y),
        )
        manifest[name] = dict(kind="decode", nc=nc, imgsz=list(imgsz), bsz=bsz, dtype=str(dtype).split(".")[-1])
        print(name, tuple(y.shape))

    for name, pk, kw in NMS_CASES:
        pred = synth_prediction(**pk)
        if name == "nms_f16_saturated":  # many scores exactly 1.0 / exactly equal -> massive ties
            pred[:, 4:] = (pred[:, 4:].float() * 6).clamp(max=1.0).to(pred.dtype)
            pred[1, 4:] = 0.0  # and one completely empty image
        stable = reference_nms(ref, pred, True, **kw)
        plain = reference_nms(ref, pred, False, **kw)
        same = all(torch.equal(a, b) for a, b in zip(stable, plain))
        if pk["dtype"] == torch.float32 and not same:
            # only exact score ties may reorder rows: same score sequence, same multiset of rows
            for a, b in zip(stable, plain):
                assert torch.equal(a[:, 4], b[:, 4]), f"{name}: score sequences differ"
                assert sorted(map(tuple, a.tolist())) == sorted(map(tuple, b.tolist())), f"{name}: row sets differ"
        rows, counts = _pack(stable)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), pred=_np(pred), rows=rows, counts=counts)
        manifest[name] = dict(kind="nms", kwargs=kw, plain_equal=bool(same), counts=counts.tolist(),
                              dtype=str(pk["dtype"]).split(".")[-1])
        print(name, counts.tolist(), "plain_equal", same)

    # crafted: IoU exactly equal to the fp32 image of the threshold and
    # an exact duplicate pair; prediction rows are (cx, cy, w, h, score) with nc = 1.
    boxes = torch.tensor([[2.5, 2.5, 5, 5], [2.5, 1.5, 5, 3], [10, 10, 20, 20], [10, 4.5, 20, 9],
                          [40, 40, 8, 8], [40, 40, 8, 8]], dtype=torch.float32)
    scores = torch.tensor([0.9, 0.8, 0.7, 0.6, 0.5, 0.5])
    pred = torch.cat((boxes, scores[:, None]), 1).t()[None].contiguous()
    for thr in (0.6, 0.45):
        nm = f"nms_f32_crafted_iou{int(thr * 100)}"
        kw = dict(conf_thres=0.25, iou_thres=thr)
        stable = reference_nms(ref, pred, True, **kw)
        plain = reference_nms(ref, pred, False, **kw)
        rows, counts = _pack(stable)
        np.savez_compressed(os.path.join(OUT, nm + ".npz"), pred=_np(pred), rows=rows, counts=counts)
        manifest[nm] = dict(kind="nms", kwargs=kw, plain_equal=bool(torch.equal(stable[0], plain[0])),
                            counts=counts.tolist(), dtype="float32")
        print(nm, counts.tolist(), rows[:, :4].tolist())

    # BASELINE config 1: the REAL 2-task CerberusDet (yolov8x_voc_obj365.yaml, random init, 105.4 M parameters) built
    # and split exactly as the reference does (utils/models_manager.py:199-213), one image, fp32, on the CPU.  Its
    # Detect heads give both the raw per-level tensors and y, i.e. a decode vector produced by the unmodified model
    # graph, and its scores (~6e-4, bias init models/yolo.py:110) give an NMS vector with > 30000 near-tied candidates.
    # 320x320 instead of 640x640 keeps the fixture small (2100 anchors).
    import copy

    import cerberusdet.models.cerberus as cerb_mod

    torch.manual_seed(0)
    model = cerb_mod.CerberusDet(task_ids=["voc", "objects365_animals"], nc=[20, 19],
                                 cfg=os.path.join(os.environ.get("CERB_REFERENCE_ROOT", "/root/reference"),
                                                  "cerberusdet/models/yolov8x_voc_obj365.yaml"), ch=3, verbose=False)
    model.sequential_split(copy.deepcopy(model.yaml["cerber"]), "cpu")
    model.eval()
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = model(torch.rand(1, 3, 320, 320))
    kw = dict(conf_thres=0.0003, iou_thres=0.6, multi_label=True)
    for task, (y, xs) in out.items():
        nm = f"model_cfg1_320_{task}"
        stable = reference_nms(ref, y, True, **kw)
        plain = reference_nms(ref, y, False, **kw)
        rows, counts = _pack(stable)
        np.savez_compressed(os.path.join(OUT, nm + ".npz"), **{f"level{i}": _np(t) for i, t in enumerate(xs)},
                            y=_np(y), pred=_np(y), rows=rows, counts=counts)
        manifest[nm] = dict(kind="model", nc=int(y.shape[1] - 4), imgsz=[320, 320], bsz=1, dtype="float32", kwargs=kw,
                            plain_equal=bool(torch.equal(stable[0], plain[0])), counts=counts.tolist())
        print(nm, tuple(y.shape), counts.tolist(), "plain_equal", manifest[nm]["plain_equal"])

    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

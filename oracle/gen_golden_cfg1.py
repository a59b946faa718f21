"""Golden vectors of BASELINE config 1 AT ITS STATED SIZE: the real 2-task CerberusDet (yolov8x_voc_obj365.yaml, VOC 20 +
Objects365-animals 19 classes, random init, 105.4 M parameters) built and split exactly as the reference does
(utils/models_manager.py:199-213), batch 1 at 640x640, fp32, on the CPU -- the unmodified model graph's Detect heads give
the raw per-level tensors and y, and the reference's own non_max_suppression the rows.  (oracle/gen_golden.py holds the
320x320 variant of the same model.)

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs the reference tree):
    python oracle/gen_golden_cfg1.py
Random-init class scores sit around 6e-4 (bias init models/yolo.py:110), below even conf 0.001, so the NMS vector uses
conf 0.0003 (> 30000 near-tied fp32 candidates per head: the max_nms cut and the top-k path at full 8400-anchor size);
the plain call at the reference's default conf 0.25 is stored too (it must return nothing).
"""
from __future__ import annotations

import copy
import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.gen_golden import OUT, _np, _pack, reference_nms  # noqa: E402
from oracle.ref_import import REFERENCE_ROOT, load_reference  # noqa: E402


def main():
    ref = load_reference()
    torch.set_num_threads(os.cpu_count())
    import cerberusdet.models.cerberus as cerb_mod

    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = cerb_mod.CerberusDet(task_ids=["voc", "objects365_animals"], nc=[20, 19],
                                     cfg=os.path.join(REFERENCE_ROOT, "cerberusdet/models/yolov8x_voc_obj365.yaml"), ch=3, verbose=False)
        model.sequential_split(copy.deepcopy(model.yaml["cerber"]), "cpu")
        model.eval()
        with torch.no_grad():
            out = model(torch.rand(1, 3, 640, 640))
    with open(os.path.join(OUT, "manifest.json")) as f:
        manifest = json.load(f)
    kw = dict(conf_thres=0.0003, iou_thres=0.6, multi_label=True)
    for task, (y, xs) in out.items():
        nm = f"model_cfg1_640_{task}"
        stable = reference_nms(ref, y, True, **kw)
        plain = reference_nms(ref, y, False, **kw)
        default = ref.general.non_max_suppression(y)  # conf 0.25 / iou 0.45: nothing passes at random init
        assert default[0].shape == (0, 6)
        rows, counts = _pack(stable)
        np.savez_compressed(os.path.join(OUT, nm + ".npz"), **{f"level{i}": _np(t) for i, t in enumerate(xs)},
                            y=_np(y), pred=_np(y), rows=rows, counts=counts)
        manifest[nm] = dict(kind="model", nc=int(y.shape[1] - 4), imgsz=[640, 640], bsz=1, dtype="float32", kwargs=kw,
                            plain_equal=bool(torch.equal(stable[0], plain[0])), counts=counts.tolist())
        print(nm, tuple(y.shape), counts.tolist(), "plain_equal", manifest[nm]["plain_equal"],
              "max score", float(y[:, 4:].max()))
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

"""Golden vectors for the TAL assigner, produced by the UNMODIFIED reference
``TaskAlignedAssigner.forward`` (cerberusdet/utils/tal.py:56-178) on the seeded inputs of ``oracle/ref_port.tal_case``.

    python oracle/gen_golden_tal.py        (build container only: needs /root/reference)

TEST INFRASTRUCTURE ONLY.  Writes tests/golden/tal_*.npz (the five output tensors; the inputs are regenerated from the
seed) and merges its entries into manifest.json.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_port as rp  # noqa: E402
from oracle.ref_import import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (seed, B, level shapes, strides, nc, max boxes per image, jitter of the predicted boxes, score dtype)
    "tal_f32_160": dict(seed=1, bs=2, level_hw=[(20, 20), (10, 10), (5, 5)], strides=[8, 16, 32], nc=7, n_gt=9, noise=4.0),
    "tal_f32_wide": dict(seed=2, bs=3, level_hw=[(16, 24), (8, 12), (4, 6)], strides=[8, 16, 32], nc=20, n_gt=14, noise=4.0),
    "tal_f32_crowded": dict(seed=3, bs=1, level_hw=[(40, 40), (20, 20), (10, 10)], strides=[8, 16, 32], nc=3, n_gt=30, noise=6.0),
    "tal_f16_scores": dict(seed=4, bs=2, level_hw=[(20, 28), (10, 14), (5, 7)], strides=[8, 16, 32], nc=12, n_gt=11, noise=3.0,
                           score_dtype=torch.float16),
}


def main():
    load_reference()
    from cerberusdet.utils.tal import TaskAlignedAssigner

    manifest_path = os.path.join(OUT, "manifest.json")
    with open(manifest_path) as f:
        manifest = json.load(f)
    for name, kw in CASES.items():
        c = rp.tal_case(**kw)
        asg = TaskAlignedAssigner(topk=10, num_classes=kw["nc"], alpha=0.5, beta=6.0)  # the loss's settings (utils/loss.py:100-105)
        labels, bboxes, scores, fg, gidx = asg(c["pd_scores"], c["pd_bboxes"], c["anc_points"], c["gt_labels"], c["gt_bboxes"], c["mask_gt"])
        np.savez_compressed(os.path.join(OUT, name + ".npz"), target_labels=labels.numpy(), target_bboxes=bboxes.numpy(),
                            target_scores=scores.numpy(), fg_mask=fg.numpy(), target_gt_idx=gidx.numpy())
        meta = {k: (v if not isinstance(v, torch.dtype) else str(v).split(".")[-1]) for k, v in kw.items()}
        manifest[name] = dict(kind="tal", **meta)
        print(name, tuple(scores.shape), "foreground anchors", int(fg.sum()))
    with open(manifest_path, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

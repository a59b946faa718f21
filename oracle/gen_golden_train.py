"""Golden vectors for the training-time sibling decode, produced by the UNMODIFIED reference:
``Loss.bbox_decode`` (cerberusdet/utils/loss.py:126-131) on anchor points from ``make_anchors``
(cerberusdet/utils/tal.py:181-193), plus the gradient torch autograd gives through that reference function.

    python oracle/gen_golden_train.py        (build container only: needs /root/reference)

TEST INFRASTRUCTURE ONLY.  Writes tests/golden/train_bbox_*.npz and merges its entries into manifest.json.
"""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = [
    # name, B, (H, W) per level, dtype, logit scale
    ("train_bbox_f32_64x96", 2, [(8, 12), (4, 6), (2, 3)], torch.float32, 3.0),
    ("train_bbox_f32_wide", 1, [(20, 20), (10, 10)], torch.float32, 12.0),
    ("train_bbox_f16_128", 3, [(16, 16), (8, 8), (4, 4)], torch.float16, 3.0),
    ("train_bbox_f16_wide", 2, [(10, 14)], torch.float16, 9.0),
]


def main():
    load_reference()
    from cerberusdet.utils.loss import Loss
    from cerberusdet.utils.tal import make_anchors

    manifest_path = os.path.join(OUT, "manifest.json")
    with open(manifest_path) as f:
        manifest = json.load(f)
    for name, bsz, shapes, dtype, scale in CASES:
        g = torch.Generator().manual_seed(sum(map(ord, name)))
        feats = [torch.zeros(bsz, 1, h, w, dtype=dtype) for h, w in shapes]
        strides = torch.tensor([8.0, 16.0, 32.0][: len(shapes)])
        anchor_points, _ = make_anchors(feats, strides, 0.5)  # [A, 2] in the feature dtype
        A = anchor_points.shape[0]
        pred = (torch.randn(bsz, A, 64, generator=g) * scale).to(dtype).requires_grad_(True)
        self = types.SimpleNamespace(use_dfl=True, proj=torch.arange(16, dtype=torch.float))
        out = Loss.bbox_decode(self, anchor_points, pred)
        assert out.dtype == dtype and tuple(out.shape) == (bsz, A, 4)
        grad_out = torch.randn(bsz, A, 4, generator=g).to(dtype)
        (grad_in,) = torch.autograd.grad(out, pred, grad_out)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), pred=pred.detach().numpy(), anchor_points=anchor_points.numpy(),
                            out=out.detach().numpy(), grad_out=grad_out.numpy(), grad_in=grad_in.numpy())
        manifest[name] = dict(kind="train_bbox", bsz=bsz, anchors=A, dtype=str(dtype).split(".")[-1])
        print(name, tuple(out.shape), float(out.float().abs().max()))
    with open(manifest_path, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

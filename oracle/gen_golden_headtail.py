"""Golden vectors for SURVEY 8f row 3 (head-tail fusion), produced by the UNMODIFIED reference ``Detect`` module
(cerberusdet/models/yolo.py:64-100) with its real conv towers: the inputs of the last 1x1 convolutions of both towers
(captured with forward pre-hooks), their parameters, and what ``Detect.forward`` returns in eval mode.

    python oracle/gen_golden_headtail.py        (build container only: needs /root/reference)

TEST INFRASTRUCTURE ONLY.  Writes tests/golden/headtail_*.npz and merges its entries into manifest.json.
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = [
    # name, nc, ch, B, (H, W) of P3, dtype
    ("headtail_f32_nc20", 20, (32, 64, 64), 2, (16, 24), torch.float32),
    ("headtail_f16_nc12", 12, (64, 64, 128), 1, (16, 16), torch.float16),
    ("headtail_f16_nc20_v8x", 20, (320, 640, 640), 1, (16, 16), torch.float16),  # yolov8x widths: c2 = 80, c3 = 320
]


def main():
    ref = load_reference()
    manifest_path = os.path.join(OUT, "manifest.json")
    with open(manifest_path) as f:
        manifest = json.load(f)
    for name, nc, ch, bsz, (h, w), dtype in CASES:
        torch.manual_seed(sum(map(ord, name)))
        m = ref.yolo.Detect(nc=nc, ch=ch)
        m.stride = torch.tensor([8.0, 16.0, 32.0])
        for pname, p in m.named_parameters():  # spread the logits (the default init gives near-constant outputs)
            if not pname.startswith("dfl."):    # (the DFL projection arange(16) is a frozen parameter: models/yolo.py:52-54)
                torch.nn.init.normal_(p, std=0.5)
        for mod in m.modules():  # BatchNorm running stats away from (0, 1) so that eval-mode BN is not the identity
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.2)
                mod.running_var.uniform_(0.5, 1.5)
        m.eval()
        if dtype == torch.float16:
            m.half()
        captured = {}

        def grab(key):
            def hook(_mod, args):
                captured[key] = args[0].detach().clone()
            return hook

        hooks = []
        for i in range(3):
            hooks.append(m.cv2[i][-1].register_forward_pre_hook(grab(("box", i))))
            hooks.append(m.cv3[i][-1].register_forward_pre_hook(grab(("cls", i))))
        feats = [torch.randn(bsz, c, h >> i, w >> i).to(dtype) for i, c in enumerate(ch)]
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            y, raw = m([f.clone() for f in feats])
        for hk in hooks:
            hk.remove()
        assert y.dtype == dtype and torch.isfinite(y.float()).all()
        arrays = {"y": y.numpy()}
        for i in range(3):
            arrays[f"box_feat{i}"] = captured[("box", i)].numpy()
            arrays[f"cls_feat{i}"] = captured[("cls", i)].numpy()
            arrays[f"box_w{i}"] = m.cv2[i][-1].weight.detach().numpy()
            arrays[f"box_b{i}"] = m.cv2[i][-1].bias.detach().numpy()
            arrays[f"cls_w{i}"] = m.cv3[i][-1].weight.detach().numpy()
            arrays[f"cls_b{i}"] = m.cv3[i][-1].bias.detach().numpy()
            arrays[f"raw{i}"] = raw[i].numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        manifest[name] = dict(kind="headtail", nc=nc, bsz=bsz, dtype=str(dtype).split(".")[-1],
                              c2=int(arrays["box_w0"].shape[1]), c3=int(arrays["cls_w0"].shape[1]))
        print(name, tuple(y.shape), "c2", arrays["box_w0"].shape[1], "c3", arrays["cls_w0"].shape[1],
              "max score", float(y[:, 4:].float().max()))
    with open(manifest_path, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

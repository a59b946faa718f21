"""The oracle restatement against the UNMODIFIED reference on fresh random inputs (beyond the committed golden vectors).
Needs the reference tree: /root/reference in the build container, its file-for-file copy oracle/_ref (oracle/make_ref.py)
anywhere else."""
import os
import sys
import types
import warnings

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_port as rp  # noqa: E402
from oracle.ref_import import reference_available  # noqa: E402

pytestmark = pytest.mark.skipif(not reference_available(), reason="needs the reference tree (/root/reference or oracle/_ref)")


@pytest.fixture(scope="module")
def ref():
    from oracle.ref_import import load_reference

    return load_reference()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("nc,shape,strides", [(7, (24, 40), (8.0, 16.0, 32.0)), (80, (16, 16), (8.0, 16.0, 32.0)),
                                              (1, (8, 12), (8.0, 16.0, 32.0))])
def test_decode_port_equals_reference_detect(ref, dtype, nc, shape, strides):
    """Reference Detect.forward eval branch (models/yolo.py:93-99) with its conv towers replaced by channel slices."""
    from oracle.gen_golden import reference_decode

    g = torch.Generator().manual_seed(nc * 1000 + shape[0])
    h, w = shape
    levels = [(torch.randn(2, 64 + nc, h >> i, w >> i, generator=g) * 4).to(dtype) for i in range(3)]
    want = reference_decode(ref, levels, nc)
    got = rp.decode_port(levels, nc, strides)
    assert got.dtype == want.dtype and torch.equal(got, want)


@pytest.mark.parametrize("kw", [
    dict(conf_thres=0.25, iou_thres=0.45),
    dict(conf_thres=0.05, iou_thres=0.6, multi_label=True, max_det=40),
    dict(conf_thres=0.1, iou_thres=0.5, classes=[0, 2], agnostic=True),
    dict(conf_thres=0.3, iou_thres=0.3, classes=[1], multi_label=True, max_det=7),
])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_nms_port_equals_reference_on_tie_free_inputs(ref, dtype, kw):
    """Reference non_max_suppression (utils/general.py:360-481), one image per call (dodges its time limit).  fp32
    random scores are tie-free, so the reference's unstable argsort defines the answer; for fp16 (ties are likely)
    the reference runs with a stable argsort, which is the canonical order of the oracle."""
    from oracle.gen_golden import reference_nms

    g = torch.Generator().manual_seed(len(str(kw)) + (1 if dtype == torch.float16 else 0))
    nc, A = 5, 900
    boxes = torch.cat((torch.rand(2, 2, A, generator=g) * 600, torch.rand(2, 2, A, generator=g) * 120 + 4), 1)
    scores = torch.rand(2, nc, A, generator=g) ** 3
    pred = torch.cat((boxes, scores), 1).to(dtype)
    want = reference_nms(ref, pred, dtype == torch.float16, **kw)
    for greedy in ("torchvision", "c"):
        got = rp.nms_port(pred, greedy=greedy, **kw)
        assert len(got) == len(want)
        for a, b in zip(got, want):
            assert tuple(a.shape) == tuple(b.shape) and torch.equal(a, b)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_bbox_decode_port_equals_reference_loss_method(ref, dtype):
    """Reference Loss.bbox_decode (utils/loss.py:126-131) and its autograd gradient."""
    from cerberusdet.utils.loss import Loss
    from cerberusdet.utils.tal import make_anchors

    g = torch.Generator().manual_seed(5)
    feats = [torch.zeros(1, 1, 12 >> i, 20 >> i, dtype=dtype) for i in range(3)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ap, _ = make_anchors(feats, torch.tensor([8.0, 16.0, 32.0]), 0.5)
    pred = (torch.randn(3, ap.shape[0], 64, generator=g) * 5).to(dtype)
    go = torch.randn(3, ap.shape[0], 4, generator=g).to(dtype)
    me = types.SimpleNamespace(use_dfl=True, proj=torch.arange(16, dtype=torch.float))
    p1, p2 = pred.clone().requires_grad_(True), pred.clone().requires_grad_(True)
    want, got = Loss.bbox_decode(me, ap, p1), rp.bbox_decode_port(ap, p2)
    assert torch.equal(got, want)
    (gw,) = torch.autograd.grad(want, p1, go)
    (gg,) = torch.autograd.grad(got, p2, go)
    assert torch.equal(gg, gw)


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_tal_assign_port_matches_reference_assigner(seed):
    """Fresh seeded inputs through the UNMODIFIED TaskAlignedAssigner.forward (utils/tal.py:56-178) and the port."""
    from oracle.ref_import import load_reference

    load_reference()
    from cerberusdet.utils.tal import TaskAlignedAssigner

    nc = 5 + seed % 4
    c = rp.tal_case(seed, 2, [(24, 16), (12, 8), (6, 4)], [8, 16, 32], nc, 8 + seed % 5)
    want = TaskAlignedAssigner(topk=10, num_classes=nc, alpha=0.5, beta=6.0)(
        c["pd_scores"], c["pd_bboxes"], c["anc_points"], c["gt_labels"], c["gt_bboxes"], c["mask_gt"])
    got = rp.tal_assign_port(**c, num_classes=nc)
    for a, b in zip(got[:5], want):
        assert a.dtype == b.dtype and torch.equal(a, b)
    # no boxes at all: the reference's early return (tal.py:89-93)
    empty = {k: (v[:, :0] if k.startswith(("gt_", "mask_")) else v) for k, v in c.items()}
    want0 = TaskAlignedAssigner(topk=10, num_classes=nc)(empty["pd_scores"], empty["pd_bboxes"], empty["anc_points"],
                                                          empty["gt_labels"], empty["gt_bboxes"], empty["mask_gt"])
    got0 = rp.tal_assign_port(**empty, num_classes=nc)
    for a, b in zip(got0[:5], want0):
        assert torch.equal(a.float(), b.float())

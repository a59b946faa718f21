"""The oracle (oracle/ref_port.py + oracle/greedy_nms.c) against the golden vectors that
oracle/gen_golden.py produced by executing the unmodified reference."""
import numpy as np
import pytest
import torch

from conftest import golden_manifest, golden_names, load_golden, split_rows
from cerberusdet_b200.synth import STRIDES
from oracle import ref_port as rp


@pytest.mark.parametrize("name", golden_names("decode"))
def test_decode_port_bit_exact(name):
    g = load_golden(name)
    nc = golden_manifest()[name]["nc"]
    levels = [torch.from_numpy(g[f"level{i}"]) for i in range(3)]
    y = rp.decode_port(levels, nc, STRIDES)
    ref = torch.from_numpy(g["y"])
    assert y.dtype == ref.dtype and y.shape == ref.shape
    assert torch.equal(y, ref)  # same ATen kernels -> bit exact on CPU


@pytest.mark.parametrize("greedy", ["torchvision", "c"])
@pytest.mark.parametrize("name", golden_names("nms"))
def test_nms_port_bit_exact(name, greedy):
    g = load_golden(name)
    kw = dict(golden_manifest()[name]["kwargs"])
    pred = torch.from_numpy(g["pred"])
    want = split_rows(g["rows"], g["counts"])
    got = rp.nms_port(pred, greedy=greedy, **kw)
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a.dtype == torch.float32 and a.shape == b.shape
        assert torch.equal(a, b)


def test_greedy_c_matches_torchvision_random():
    import torchvision

    gen = torch.Generator().manual_seed(5)
    for n in (0, 1, 2, 65, 700, 3000):
        c = torch.rand(n, 2, generator=gen) * 200
        wh = torch.rand(n, 2, generator=gen) * 60
        boxes = torch.cat((c - wh / 2, c + wh / 2), 1)
        boxes[n // 2 :] = boxes[n // 2 :].round()  # integer boxes: many exact-equality IoUs
        scores = torch.rand(n, generator=gen).sort(descending=True).values
        for thr in (0.0, 0.1, 0.45, 0.5, 0.6, 0.7, 1.0):
            assert torch.equal(torchvision.ops.nms(boxes, scores, thr), rp.greedy_nms_c(boxes, thr))


def test_threshold_rounding_rules():
    # conf threshold is compared in the tensor dtype (utils/general.py:411)
    h = torch.tensor([0.30005], dtype=torch.float16)
    pred = torch.zeros(1, 5, 1, dtype=torch.float16)
    pred[0, :4, 0] = torch.tensor([10, 10, 4, 4], dtype=torch.float16)
    pred[0, 4, 0] = h
    assert rp.nms_port(pred, conf_thres=0.3)[0].shape[0] == 0  # half(0.3) == 0.30005, not >
    f = torch.tensor(np.float32(0.3))
    pred32 = pred.float()
    pred32[0, 4, 0] = f
    assert rp.nms_port(pred32, conf_thres=0.3)[0].shape[0] == 0
    pred32[0, 4, 0] = torch.tensor(np.nextafter(np.float32(0.3), np.float32(1)))
    assert rp.nms_port(pred32, conf_thres=0.3)[0].shape[0] == 1


def test_asserts_like_reference():
    pred = torch.zeros(1, 6, 4)
    with pytest.raises(AssertionError):
        rp.nms_port(pred, conf_thres=1.5)
    with pytest.raises(AssertionError):
        rp.nms_port(pred, iou_thres=-0.1)


@pytest.mark.parametrize("name", golden_names("model"))
def test_real_model_config1_vectors(name):
    """BASELINE config 1: vectors produced by the real 2-task CerberusDet model graph (random init) on the CPU."""
    g = load_golden(name)
    meta = golden_manifest()[name]
    levels = [torch.from_numpy(g[f"level{i}"]) for i in range(3)]
    y = rp.decode_port(levels, meta["nc"], STRIDES)
    assert torch.equal(y, torch.from_numpy(g["y"]))
    for greedy in ("torchvision", "c"):
        got = rp.nms_port(y, greedy=greedy, **meta["kwargs"])
        want = split_rows(g["rows"], g["counts"])
        assert len(got) == 1 and torch.equal(got[0], want[0])


# ------------------------------------------------------------------ training-time sibling decode
@pytest.mark.parametrize("name", golden_names("train_bbox"))
def test_bbox_decode_port_bit_exact(name):
    """The oracle restatement of Loss.bbox_decode (utils/loss.py:126-131) against vectors the unmodified reference
    produced (oracle/gen_golden_train.py): forward and the autograd gradient, bit for bit (same ATen kernels)."""
    from oracle import ref_port as rp

    g = load_golden(name)
    pred = torch.from_numpy(g["pred"]).requires_grad_(True)
    out = rp.bbox_decode_port(torch.from_numpy(g["anchor_points"]), pred)
    assert torch.equal(out.detach(), torch.from_numpy(g["out"]))
    (grad_in,) = torch.autograd.grad(out, pred, torch.from_numpy(g["grad_out"]))
    assert torch.equal(grad_in, torch.from_numpy(g["grad_in"]))


# ------------------------------------------------------------------ head tail: the checker, ahead of the kernel
@pytest.mark.parametrize("name", golden_names("headtail"))
def test_head_tail_port_bit_exact(name):
    """Last 1x1 convs of both towers + concat + decode, restated in the oracle, against what the unmodified reference
    Detect module (real conv towers, eval mode) returned for the captured inputs of those convs
    (oracle/gen_golden_headtail.py)."""
    from oracle import ref_port as rp

    g = load_golden(name)
    t = lambda k: torch.from_numpy(g[k])  # noqa: E731
    y, raw = rp.head_tail_port([t(f"box_feat{i}") for i in range(3)], [t(f"cls_feat{i}") for i in range(3)],
                               [t(f"box_w{i}") for i in range(3)], [t(f"box_b{i}") for i in range(3)],
                               [t(f"cls_w{i}") for i in range(3)], [t(f"cls_b{i}") for i in range(3)], (8.0, 16.0, 32.0))
    for i in range(3):
        assert torch.equal(raw[i], t(f"raw{i}")), f"raw level {i}"
    assert torch.equal(y, t("y"))


def _tal_case_of(name):
    meta = dict(golden_manifest()[name])
    meta.pop("kind")
    if "score_dtype" in meta:
        meta["score_dtype"] = getattr(torch, meta["score_dtype"])
    return rp.tal_case(**meta), meta


@pytest.mark.parametrize("name", golden_names("tal"))
def test_tal_assign_port_bit_exact(name):
    """oracle/ref_port.tal_assign_port against what the unmodified TaskAlignedAssigner.forward returned on the same seeded
    inputs (oracle/gen_golden_tal.py): all five tensors, bit for bit."""
    g = load_golden(name)
    c, meta = _tal_case_of(name)
    labels, bboxes, scores, fg, gidx, ambiguous = rp.tal_assign_port(**c, num_classes=meta["nc"])
    assert torch.equal(labels, torch.from_numpy(g["target_labels"]))
    assert torch.equal(bboxes, torch.from_numpy(g["target_bboxes"]))
    assert torch.equal(scores, torch.from_numpy(g["target_scores"]))
    assert torch.equal(fg, torch.from_numpy(g["fg_mask"]))
    assert torch.equal(gidx, torch.from_numpy(g["target_gt_idx"]))
    assert fg.any() and (fg.sum(1) > 0).all()

"""GPU parity of the TAL assigner kernels (``cerb_tal_assign``: TaskAlignedAssigner.forward,
cerberusdet/utils/tal.py:56-178, in three launches) against the golden vectors of the unmodified reference, the oracle
port on seeded inputs, and -- where the reference tree is present (oracle/_ref on the GPU box) -- the reference's own
assigner running on the same GPU, plus the ``patch.install(train=True)`` wiring.

Bar: indices and masks (target_labels, fg_mask, target_gt_idx, target_bboxes) bit-exact; target_scores bit-exact against
the reference ON THE SAME GPU (same libdevice atanf / powf as the ATen kernels), 1e-5 relative against CPU-generated
goldens (the CPU's atan / pow differ from the GPU's in the last bits).  Ground-truth boxes whose top-k is ambiguous in the
reference (fewer than topk positive-metric anchors and zero-metric anchors inside: ``tal_assign_port``'s last return
value) are compared against the port only, which shares the kernels' tie rule."""
import os
import sys
import warnings

import pytest
import torch

from conftest import golden_manifest, golden_names, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_port as rp  # noqa: E402
from oracle.ref_import import reference_available  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from cerberusdet_b200 import _lib, ops as o

    _lib.load()
    return o


def _cuda(c):
    return {k: v.cuda() for k, v in c.items()}


def _golden_case(name):
    meta = dict(golden_manifest()[name])
    meta.pop("kind")
    if "score_dtype" in meta:
        meta["score_dtype"] = getattr(torch, meta["score_dtype"])
    return rp.tal_case(**meta), meta


def _check(got, want, scores_exact, images=None):
    names = ("target_labels", "target_bboxes", "target_scores", "fg_mask", "target_gt_idx")
    for n, a, b in zip(names, got, want):
        a, b = a.cpu(), b.cpu()
        if images is not None:
            a, b = a[images], b[images]
        assert a.dtype == b.dtype and a.shape == b.shape, n
        if n == "target_scores" and not scores_exact:
            assert ((a > 0) == (b > 0)).all(), n
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-9), n
        else:
            assert torch.equal(a, b), n


@pytest.mark.parametrize("name", golden_names("tal"))
def test_tal_assign_golden(ops, name):
    g = load_golden(name)
    c, meta = _golden_case(name)
    got = ops.tal_assign(**_cuda(c), topk=10, num_classes=meta["nc"], alpha=0.5, beta=6.0)
    want = [torch.from_numpy(g[k]) for k in ("target_labels", "target_bboxes", "target_scores", "fg_mask", "target_gt_idx")]
    _check(got, want, scores_exact=False)


@pytest.mark.parametrize("seed,bs,level_hw,nc,n_gt,dtype", [
    (21, 4, [(80, 80), (40, 40), (20, 20)], 20, 20, torch.float32),   # the training shape: 640x640, 8400 anchors
    (22, 2, [(80, 80), (40, 40), (20, 20)], 12, 60, torch.float16),   # crowded, half scores
    (23, 3, [(12, 20), (6, 10), (3, 5)], 3, 4, torch.float32),        # fewer than topk anchors inside most boxes
    (24, 1, [(4, 4)], 2, 1, torch.float32),                           # one level, 16 anchors, one box
])
def test_tal_assign_vs_oracle_port_on_the_same_gpu(ops, seed, bs, level_hw, nc, n_gt, dtype):
    """The port shares the kernels' tie rule, so every index and mask must agree bit for bit, ambiguous boxes included."""
    c = _cuda(rp.tal_case(seed, bs, level_hw, [8, 16, 32][: len(level_hw)], nc, n_gt, score_dtype=dtype))
    got = ops.tal_assign(**c, topk=10, num_classes=nc)
    # (the port indexes with CPU aranges: it runs on the CPU copies and the scores are compared with the CPU tolerance;
    # the exact comparison of the scores is the next test, against the reference itself on this GPU)
    want = rp.tal_assign_port(**{k: v.cpu() for k, v in c.items()}, num_classes=nc)
    torch.cuda.synchronize()
    _check(got, want[:5], scores_exact=False)
    assert got[3].any()


@pytest.mark.skipif(not reference_available(), reason="needs the reference tree (/root/reference or oracle/_ref)")
@pytest.mark.parametrize("seed,dtype", [(31, torch.float32), (32, torch.float16), (33, torch.float32)])
def test_tal_assign_equals_the_reference_assigner_on_this_gpu(ops, seed, dtype):
    from oracle.ref_import import load_reference

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        load_reference()
    from cerberusdet.utils.tal import TaskAlignedAssigner

    nc, bs = 20, 8
    c = rp.tal_case(seed, bs, [(80, 80), (40, 40), (20, 20)], [8, 16, 32], nc, 24, score_dtype=dtype)
    ambiguous = rp.tal_assign_port(**c, num_classes=nc)[5]
    clean = [i for i in range(bs) if not bool(ambiguous[i].any())]
    assert len(clean) >= bs // 2, "the seeded case should be unambiguous for most images"
    d = _cuda(c)
    asg = TaskAlignedAssigner(topk=10, num_classes=nc, alpha=0.5, beta=6.0).cuda()
    fwd = getattr(TaskAlignedAssigner.forward, "_cerb_reference", TaskAlignedAssigner.forward)
    want = fwd(asg, d["pd_scores"], d["pd_bboxes"], d["anc_points"], d["gt_labels"], d["gt_bboxes"], d["mask_gt"])
    got = ops.tal_assign(**d, topk=10, num_classes=nc)
    torch.cuda.synchronize()
    _check(got, want, scores_exact=True, images=clean)


@pytest.mark.skipif(not reference_available(), reason="needs the reference tree (/root/reference or oracle/_ref)")
def test_install_train_rebinds_the_assigner_and_make_anchors(ops):
    from oracle.ref_import import load_reference

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        load_reference()
    import cerberusdet.utils.loss as loss_mod
    from cerberusdet.utils.tal import TaskAlignedAssigner

    from cerberusdet_b200 import patch

    patch.uninstall()
    nc = 7
    c = rp.tal_case(41, 2, [(20, 20), (10, 10), (5, 5)], [8, 16, 32], nc, 9)
    asg = TaskAlignedAssigner(topk=10, num_classes=nc, alpha=0.5, beta=6.0)
    want_cpu = asg(c["pd_scores"], c["pd_bboxes"], c["anc_points"], c["gt_labels"], c["gt_bboxes"], c["mask_gt"])
    feats = [torch.zeros(2, 1, h, w, device="cuda") for h, w in [(20, 20), (10, 10), (5, 5)]]
    strides = torch.tensor([8.0, 16.0, 32.0])
    want_anchors = loss_mod.make_anchors(feats, strides, 0.5)
    try:
        info = patch.install(train=True)
        assert "cerberusdet.utils.tal.TaskAlignedAssigner.forward" in info["patched"]
        assert "cerberusdet.utils.loss.make_anchors" in info["patched"]
        calls = []
        real = ops.tal_assign
        ops.tal_assign = lambda *a, **k: (calls.append(1), real(*a, **k))[1]
        try:
            d = _cuda(c)
            got = asg(d["pd_scores"], d["pd_bboxes"], d["anc_points"], d["gt_labels"], d["gt_bboxes"], d["mask_gt"])
            assert calls, "the kernel path was not taken"
            _check(got, want_cpu, scores_exact=False)
            n_calls = len(calls)
            got_cpu = asg(c["pd_scores"], c["pd_bboxes"], c["anc_points"], c["gt_labels"], c["gt_bboxes"], c["mask_gt"])  # CPU: reference code
            assert len(calls) == n_calls
            _check(got_cpu, want_cpu, scores_exact=True)
            empty = asg(d["pd_scores"], d["pd_bboxes"], d["anc_points"], d["gt_labels"][:, :0], d["gt_bboxes"][:, :0], d["mask_gt"][:, :0])
            assert len(calls) == n_calls and not bool(empty[3].any())  # no boxes: the reference's early return
        finally:
            ops.tal_assign = real
        a1 = loss_mod.make_anchors(feats, strides, 0.5)
        a2 = loss_mod.make_anchors(feats, strides, 0.5)
        assert a1[0] is a2[0] and torch.equal(a1[0], want_anchors[0]) and torch.equal(a1[1], want_anchors[1])
    finally:
        patch.uninstall()
    assert not hasattr(TaskAlignedAssigner.forward, "_cerb_reference")

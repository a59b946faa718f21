"""GPU tests of the drop-in on the REAL reference: ``patch.install()`` against the unmodified ``Detect``,
``non_max_suppression`` and ``CerberusDetInference`` with CUDA tensors, including the real 2-task CerberusDet model
(yolov8x_voc_obj365.yaml, BASELINE config 1) at 640x640.

The reference tree is ``/root/reference`` in the build container and its file-for-file copy ``oracle/_ref``
(``oracle/make_ref.py``, git-ignored, shipped with the gpurun snapshot) on the GPU box.  Every comparison is
"the reference's own code on this GPU" against "the same objects after ``patch.install()``".
"""
import copy
import json
import os
import sys
import warnings

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import REFERENCE_ROOT, reference_available  # noqa: E402

from cerberusdet_b200.synth import STRIDES, synth_prediction  # noqa: E402
from tol import check_decode  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not reference_available(), reason="needs the reference tree (/root/reference or oracle/_ref)")]


@pytest.fixture(scope="module")
def ref():
    from oracle.ref_import import load_reference

    os.environ.setdefault("TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD", "1")  # attempt_load unpickles whole modules (torch >= 2.6)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return load_reference()


@pytest.fixture()
def patch(ref):
    from cerberusdet_b200 import patch as p

    p.uninstall()
    yield p
    p.uninstall()


def _real_detect(ref, nc, ch, half):
    torch.manual_seed(nc)
    m = ref.yolo.Detect(nc=nc, ch=ch)
    m.stride = torch.tensor(STRIDES)
    m.bias_init()
    for seq in m.cv3:  # spread the class logits like a trained head's
        torch.nn.init.normal_(seq[-1].weight, std=0.5)
    m = m.cuda().eval()
    return m.half() if half else m


@pytest.mark.parametrize("half", [False, True])
def test_real_detect_forward_patched_vs_reference(ref, patch, half):
    """The reference's own Detect (real conv towers, cuDNN) on CUDA: after install() the same instance returns the same
    raw levels bit for bit and y within the north_star tolerance; attributes other code reads stay populated."""
    m = _real_detect(ref, 20, (64, 128, 256), half)
    dt = torch.float16 if half else torch.float32
    g = torch.Generator().manual_seed(5)
    feats = [torch.randn(3, c, 640 // int(s), 640 // int(s), generator=g).to("cuda", dt) for c, s in zip((64, 128, 256), STRIDES)]
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        y_ref, x_ref = m([f.clone() for f in feats])
        info = patch.install()
        assert "cerberusdet.models.yolo.Detect.forward" in info["patched"]
        m.shape = None  # new anchor-cache miss on the patched path too
        y, x = m([f.clone() for f in feats])
    assert y.dtype == y_ref.dtype and y.shape == y_ref.shape == (3, 24, 8400)
    for a, b in zip(x, x_ref):
        assert torch.equal(a, b)
    ok, msg = check_decode(y, y_ref, [t.shape[2:] for t in x], STRIDES, 20)
    assert ok, msg
    assert tuple(m.anchors.shape) == (2, 8400) and tuple(m.strides.shape) == (1, 8400) and m.shape == feats[0].shape
    # inference_mode tensors carry no version counter: the patched forward must still work (no score summary then)
    with torch.inference_mode(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        y_inf, _ = m([f.clone() for f in feats])
        assert torch.equal(y_inf, y)
        out = ref.general.non_max_suppression(y_inf, 0.25, 0.45)
        want = ref.general.non_max_suppression(y, 0.25, 0.45)
    assert all(torch.equal(a, b) for a, b in zip(out, want))


@pytest.mark.parametrize("kw", [dict(conf_thres=0.25, iou_thres=0.45), dict(conf_thres=0.001, iou_thres=0.6, multi_label=True),
                                dict(conf_thres=0.1, iou_thres=0.5, agnostic=True, max_det=50)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_real_nms_patched_vs_reference_cuda(ref, patch, dtype, kw):
    """The reference non_max_suppression on CUDA (torchvision's CUDA kernel) against the patched name, same CUDA input.
    Scores are made distinct (the reference's argsort is unstable on ties)."""
    pred = synth_prediction(2, 7, 1500, seed=23, dtype=torch.float32, regime="clusters")
    pred[:, 4:] += torch.arange(pred[:, 4:].numel()).reshape(pred[:, 4:].shape) * 1e-7
    if dtype == torch.float16:
        pred = pred.half()
        s = pred[:, 4:].float()
        flat = s.flatten()
        _, inv, cnt = torch.unique(flat, return_inverse=True, return_counts=True)
        flat[cnt[inv] > 1] = 0.0  # drop exactly equal half scores: tie order is unspecified in the reference
        pred[:, 4:] = flat.reshape(s.shape).half()
    pred = pred.cuda()
    want = [ref.general.non_max_suppression(pred[i:i + 1], **kw)[0] for i in range(pred.shape[0])]  # one image per call (time limit)
    patch.install()
    got = ref.general.non_max_suppression(pred, **kw)
    assert hasattr(ref.general.non_max_suppression, "_cerb_reference")
    for a, b in zip(got, want):
        assert a.dtype == b.dtype == torch.float32 and a.device == b.device
        assert torch.equal(a, b), f"{kw}: {a.shape} vs {b.shape}"


def test_torchvision_cuda_threshold_convention_recorded(ref):
    """record what torchvision's CUDA kernel does on this box at exact-equality IoU.  The
    kernels follow the CPU convention ((double)ovr > thr); the record goes to gpurun_out/ and is copied into profiles/."""
    import torchvision

    from cerberusdet_b200.nms import non_max_suppression

    rec = {"torchvision": torchvision.__version__}
    for name, boxes, thr in [("iou60", [[0., 0, 5, 5], [0, 0, 5, 3]], 0.6), ("iou45", [[0., 0, 20, 20], [0, 0, 20, 9]], 0.45)]:
        b = torch.tensor(boxes)
        s = torch.tensor([0.9, 0.8])
        cpu = torchvision.ops.nms(b, s, thr).tolist()
        cuda = torchvision.ops.nms(b.cuda(), s.cuda(), thr).tolist()
        xywh = torch.stack(((b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2, b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]), 1)
        pred = torch.cat((xywh, s[:, None]), 1).t()[None].contiguous().cuda()
        ours = non_max_suppression(pred, 0.25, thr)[0].shape[0]
        rec[name] = {"torchvision_cpu_keep": cpu, "torchvision_cuda_keep": cuda, "cerb_kept": ours}
        assert ours == len(cpu), f"{name}: the kernel must follow the CPU convention"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "torchvision_cuda_threshold.json"), "w") as f:
        json.dump(rec, f, indent=1)


def _build_real_model(ref, tmp_path):
    """The real 2-task CerberusDet of BASELINE config 1, built and split as the reference does
    (utils/models_manager.py:199-213), heads re-initialised so that scores spread like a trained model's, saved as a
    checkpoint ``attempt_load`` reads."""
    import cerberusdet.models.cerberus as cerb_mod

    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = cerb_mod.CerberusDet(task_ids=["voc", "objects365_animals"], nc=[20, 19],
                                     cfg=os.path.join(REFERENCE_ROOT, "cerberusdet/models/yolov8x_voc_obj365.yaml"), ch=3, verbose=False)
        model.sequential_split(copy.deepcopy(model.yaml["cerber"]), "cpu")
    # random init leaves the activations ~1e-5 deep in the network (every score = sigmoid(bias)): give the BatchNorms
    # the statistics of one random batch so that activations are O(1) everywhere, then spread the class logits
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.momentum = 1.0
    model.train()
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model(torch.rand(2, 3, 256, 256, generator=torch.Generator().manual_seed(7)))
    model.eval()
    g = torch.Generator().manual_seed(1)
    for m in model.modules():
        if type(m).__name__ == "Detect":
            for seq in m.cv3:
                seq[-1].weight.data.normal_(0.0, 0.004, generator=g)
                seq[-1].bias.data.fill_(-5.0)
    model.names = {"voc": [f"voc{i}" for i in range(20)], "objects365_animals": [f"ani{i}" for i in range(19)]}
    path = os.path.join(tmp_path, "cerber_cfg1.pt")
    torch.save({"model": model}, path)
    return path


def _match(a, b, px=1.01, rel=2e-3):
    """Detections of one image: same labels/tasks, scores and boxes within tolerance, order-insensitive.
    Returns (fraction of a's detections found in b, message)."""
    left = list(b)
    found = 0
    miss = ""
    for d in a:
        hit = None
        for k, e in enumerate(left):
            if (d["label"] == e["label"] and d["task"] == e["task"] and d["label_name"] == e["label_name"]
                    and abs(d["score"] - e["score"]) <= rel * max(abs(e["score"]), 1e-3)
                    and all(abs(p - q) <= px for p, q in zip(d["box"], e["box"]))):
                hit = k
                break
        if hit is None:
            miss = f"unmatched {d}"
        else:
            found += 1
            left.pop(hit)
    return found / max(len(a), len(b), 1), f"{found} of {len(a)} / {len(b)} matched; {miss}"


@pytest.mark.parametrize("half,fuse", [(False, False), (True, False), (True, True)])
def test_real_cerberusdet_inference_640_patched_vs_reference(ref, patch, tmp_path, half, fuse, monkeypatch):
    """BASELINE config 1 end to end on the GPU: the reference CerberusDetInference (attempt_load, real model graph,
    per-task reference NMS, host cross-task tail) against the class install() rebinds, at 640x640.  ``fuse``: the heads'
    last 1x1 convolutions run inside the tcgen05 head-tail kernel instead of cuDNN."""
    import cerberusdet.cerberusdet_inference as inf_mod

    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", os.environ.get("CUDA_VISIBLE_DEVICES", "0"))  # select_device() rewrites it
    path = _build_real_model(ref, str(tmp_path))
    torch.manual_seed(3)
    img = torch.rand(2, 3, 640, 640, device="cuda")
    img = img.half() if half else img
    kw = dict(conf_thres=0.05, iou_thres=0.45, iou_thres_between_tasks=0.8, half=half, img_size=640)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        eng_ref = inf_mod.CerberusDetInference(path, device="0", **kw)
        assert type(eng_ref).__module__ == "cerberusdet.cerberusdet_inference"
        want = eng_ref.predict(img, original_shape=[(480, 640), (640, 427)], max_det=300)
        want_raw = eng_ref.predict(img, max_det=300)
        del eng_ref
        info = patch.install(import_all=True)
        assert "cerberusdet.cerberusdet_inference.CerberusDetInference" in info["patched"]
        eng = inf_mod.CerberusDetInference(path, device="0", **kw)
        assert type(eng).__module__ == "cerberusdet_b200.inference"
        eng.fuse_head_tail = fuse
        import cerberusdet_b200.inference as our_inf

        fused_calls = []
        real_head_tail = our_inf.head_tail
        monkeypatch.setattr(our_inf, "head_tail", lambda *a: (fused_calls.append(1), real_head_tail(*a))[1])
        got = eng.predict(img, original_shape=[(480, 640), (640, 427)], max_det=300)
        got_raw = eng.predict(img, max_det=300)
        assert bool(fused_calls) == fuse, "the fused head-tail kernel ran exactly when asked for"
    assert len(got) == len(want) == 2 and sum(len(x) for x in want) > 100, [len(x) for x in want]
    # fp32: every detection matches.  fp16: the network's half scores tie often and the reference's argsort
    # (utils/general.py:459) is unstable, so which of two equal-score overlapping boxes survives is unspecified there
    need = 0.9 if half else 1.0
    rel = 1.7e-2 if fuse else 2e-3  # fused: conv outputs may differ from cuDNN's by one half-ulp of a logit
    for i in range(2):
        frac, msg = _match(got[i], want[i], rel=rel)
        assert frac >= need, f"image {i}: {msg}"
        frac, msg = _match(got_raw[i], want_raw[i], rel=rel)
        assert frac >= need, f"image {i} (no rescale): {msg}"
    assert eng.stride == 32 and set(eng.names) == {"voc", "objects365_animals"}
    # the patched model's heads return the reference layout when called directly
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = eng.model(img)
    assert set(out) == {"voc", "objects365_animals"}
    assert tuple(out["voc"][0].shape) == (2, 24, 8400) and tuple(out["objects365_animals"][0].shape) == (2, 23, 8400)
    assert [tuple(t.shape[1:]) for t in out["voc"][1]] == [(84, 80, 80), (84, 40, 40), (84, 20, 20)]


@pytest.mark.parametrize("amp", [False, True])
def test_train_patch_bbox_decode_under_autocast(ref, patch, amp):
    """the reference trainer computes the loss under amp.autocast (trainers/averaging.py:158); the patched
    Loss.bbox_decode must take the kernel path there and agree with the reference's autocast result."""
    import types

    import cerberusdet.utils.loss as loss_mod
    from tol import check_bbox_decode

    me = types.SimpleNamespace(use_dfl=True, proj=torch.arange(16, dtype=torch.float, device="cuda"))
    dt = torch.float16 if amp else torch.float32
    g = torch.Generator().manual_seed(9)
    ap = (torch.rand(600, 2, generator=g) * 40).to("cuda", dt)
    pred = (torch.randn(2, 600, 64, generator=g) * 3).to("cuda", dt)
    go = torch.randn(2, 600, 4, generator=g).to("cuda", dt)
    orig = loss_mod.Loss.bbox_decode
    with torch.autocast("cuda", enabled=amp):
        p0 = pred.clone().requires_grad_(True)
        want = orig(me, ap, p0)
        (gw,) = torch.autograd.grad(want, p0, go.to(want.dtype))
    assert patch.install() and "cerberusdet.utils.loss.Loss.bbox_decode" in patch.install(train=True)["patched"]  # second call adds it
    calls = []
    from cerberusdet_b200 import ops

    real = ops.bbox_decode
    ops.bbox_decode = lambda a, p: (calls.append(1), real(a, p))[1]
    try:
        with torch.autocast("cuda", enabled=amp):
            p1 = pred.clone().requires_grad_(True)
            got = loss_mod.Loss.bbox_decode(me, ap, p1)
            (gg,) = torch.autograd.grad(got, p1, go.to(got.dtype))
    finally:
        ops.bbox_decode = real
    assert calls, "the kernel path was not taken"
    assert got.dtype == want.dtype
    ok, msg = check_bbox_decode(got.detach(), want.detach(), gmax=56.0)
    assert ok, msg
    if not amp:  # (under autocast the reference's softmax backward runs in fp32 on unrounded probabilities)
        ok, msg = check_bbox_decode(gg, gw, grad=True, grad_out=go)
        assert ok, msg
    else:
        assert torch.allclose(gg.float(), gw.float(), rtol=2e-2, atol=2e-2 * float(go.abs().max()) * 16)
    # mixed dtypes keep the reference's promotion (fp16 distances, fp32 anchors -> fp32 corners)
    if amp:
        with torch.autocast("cuda"):
            mixed = loss_mod.Loss.bbox_decode(me, ap.float(), pred)
            assert mixed.dtype == orig(me, ap.float(), pred).dtype


def test_val_patch_process_batch_matches_reference(ref, patch):
    """SURVEY 8f-2 wiring: install(val=True) rebinds cerberusdet.val.process_batch; CUDA inputs give the reference's
    matrix (inputs without exactly equal IoUs: the reference's numpy sort is unstable there)."""
    import cerberusdet.val as val_mod

    orig = val_mod.process_batch
    iouv = torch.linspace(0.5, 0.95, 10, device="cuda")
    cases = []
    g = torch.Generator().manual_seed(4)
    for n, m in [(300, 17), (40, 3), (5, 60), (1, 1)]:
        lab_xy = torch.rand(m, 2, generator=g) * 500
        lab_wh = 20 + torch.rand(m, 2, generator=g) * 120
        labels = torch.cat((torch.randint(0, 4, (m, 1), generator=g).float(), lab_xy, lab_xy + lab_wh), 1)
        pick = torch.randint(0, m, (n,), generator=g)
        det_box = labels[pick, 1:] + torch.randn(n, 4, generator=g) * 6
        dets = torch.cat((det_box, torch.rand(n, 1, generator=g), labels[pick, :1].clone()), 1)
        dets[::7, 5] = (dets[::7, 5] + 1) % 4
        cases.append((dets.cuda(), labels.cuda()))
    want = [orig(d, l, iouv) for d, l in cases]
    info = patch.install(val=True)
    assert "cerberusdet.val.process_batch" in info["patched"] and "cerberusdet.val.non_max_suppression" in info["patched"]
    for (d, l), w in zip(cases, want):
        got = val_mod.process_batch(d, l, iouv)
        assert got.dtype == torch.bool and got.device == d.device and torch.equal(got, w)
    assert torch.equal(val_mod.process_batch(cases[0][0].cpu(), cases[0][1].cpu(), iouv.cpu()), want[0].cpu())  # CPU: reference code
    patch.uninstall()
    assert val_mod.process_batch is orig


def test_validation_batch_statistics_equals_the_reference_loop(ref):
    """the "Statistics per image" block of the reference validation loop (cerberusdet/val.py:321-357) executed
    with the reference's OWN functions on CUDA tensors, image by image, against ``val_stats.validation_batch_statistics``
    (batched rescale + one cerb_val_match launch): the tuples the loop appends to ``stats``, bit for bit."""
    import cerberusdet.val as val_mod
    from cerberusdet.utils.general import scale_boxes, xywh2xyxy

    from cerberusdet_b200 import ops, val_stats

    process_batch = getattr(val_mod.process_batch, "_cerb_reference", val_mod.process_batch)
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(21)
    B, max_det, nc = 6, 300, 4
    img = torch.zeros(B, 3, 384, 640, device=dev)
    ori_shape = [(480, 640), (720, 1280), (384, 640), (1000, 700), (333, 500), (640, 640)]
    ratio_pad = []
    for (h0, w0) in ori_shape:  # what the reference's letterbox records: ((ratio_h, ratio_w), (pad_w, pad_h))
        r = min(384 / h0, 640 / w0)
        ratio_pad.append(((r, r), ((640 - w0 * r) / 2, (384 - h0 * r) / 2)))
    n_lab = [7, 0, 25, 3, 0, 12]
    n_det = [300, 40, 0, 5, 0, 120]
    cls_l, box_l, idx_l = [], [], []
    dets = torch.zeros(B, max_det, 6)
    for i in range(B):
        cxy = 0.15 + 0.7 * torch.rand(n_lab[i], 2, generator=g)
        wh = 0.05 + 0.25 * torch.rand(n_lab[i], 2, generator=g)
        c = torch.randint(0, nc, (n_lab[i], 1), generator=g).float()
        cls_l.append(c); box_l.append(torch.cat((cxy, wh), 1)); idx_l.append(torch.full((n_lab[i],), float(i)))
        n = n_det[i]
        if n:
            if n_lab[i]:
                pick = torch.randint(0, n_lab[i], (n,), generator=g)
                lab_xyxy = torch.cat((cxy - wh / 2, cxy + wh / 2), 1)[pick] * torch.tensor([640.0, 384.0, 640.0, 384.0])
                boxes = lab_xyxy + torch.randn(n, 4, generator=g) * 6
                pc = c[pick, 0].clone()
                pc[::5] = (pc[::5] + 1) % nc
            else:
                boxes = torch.rand(n, 4, generator=g) * 300
                boxes[:, 2:] += boxes[:, :2]
                pc = torch.randint(0, nc, (n,), generator=g).float()
            conf = torch.sort(torch.rand(n, generator=g), descending=True).values
            dets[i, :n] = torch.cat((boxes, conf[:, None], pc[:, None]), 1)
    batch = {"img": img, "batch_idx": torch.cat(idx_l), "cls": torch.cat(cls_l), "bboxes": torch.cat(box_l),
             "ori_shape": ori_shape, "ratio_pad": ratio_pad}
    batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    dets, counts = dets.to(dev), torch.tensor(n_det, dtype=torch.int32, device=dev)
    iouv = torch.linspace(0.5, 0.95, 10, device=dev)
    niou = iouv.numel()

    # ---- the reference loop body (val.py:321-357), verbatim in structure, with the reference's own functions
    want = []
    out = [dets[i, : n_det[i]].clone() for i in range(B)]
    for si, pred in enumerate(out):
        idx = batch["batch_idx"] == si
        cls, bbox = batch["cls"][idx], batch["bboxes"][idx]
        nl, npr = cls.shape[0], pred.shape[0]
        shape = batch["ori_shape"][si]
        correct_bboxes = torch.zeros(npr, niou, dtype=torch.bool, device=dev)
        if npr == 0:
            if nl:
                want.append((correct_bboxes, *torch.zeros((2, 0), device=dev), cls.squeeze(-1)))
            continue
        predn = pred.clone()
        scale_boxes(batch["img"][si].shape[1:], predn[:, :4], shape, ratio_pad=batch["ratio_pad"][si])
        if nl:
            height, width = batch["img"].shape[2:]
            tbox = xywh2xyxy(bbox) * torch.tensor((width, height, width, height), device=dev)
            scale_boxes(batch["img"][si].shape[1:], tbox, shape, ratio_pad=batch["ratio_pad"][si])
            labelsn = torch.cat((cls, tbox), 1)
            correct_bboxes = process_batch(predn, labelsn, iouv)
        want.append((correct_bboxes, pred[:, 4], pred[:, 5], cls.squeeze(-1)))

    got = val_stats.validation_batch_statistics(dets, counts, batch, iouv)
    assert len(got) == len(want) == 5  # image 4 has neither labels nor predictions: nothing appended (val.py:333-338)
    for a, b in zip(got, want):
        assert len(a) == len(b) == 4
        for x, y in zip(a, b):
            assert x.shape == y.shape and x.dtype == y.dtype and x.device == y.device and torch.equal(x, y)
    assert any(bool(t[0].any()) for t in want) and any(not bool(t[0].all()) for t in want)
    # ratio_pad = None (the reference then derives gain and pad from the shapes): same function, same result
    batch2 = dict(batch, ratio_pad=None)
    got2 = val_stats.validation_batch_statistics(dets, counts, batch2, iouv)
    want2 = []
    for si in range(B):
        if n_det[si] == 0:
            continue
        predn = dets[si, : n_det[si]].clone()
        scale_boxes(batch["img"][si].shape[1:], predn[:, :4], ori_shape[si])
        want2.append(predn)
    # (the statistics tuples do not carry the boxes; compare the rescale itself through the batched helper)
    inv_gain, px, py, w0, h0 = val_stats._scale_params((384, 640), ori_shape, [None] * B, dev)
    scaled = val_stats._scale_boxes_batched(dets[..., :4], inv_gain, px, py, w0, h0)
    k = 0
    for si in range(B):
        if n_det[si]:
            assert torch.equal(scaled[si, : n_det[si]], want2[k][:, :4])
            k += 1
    assert len(got2) == len(got)

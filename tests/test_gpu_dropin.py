"""GPU tests of the drop-in surfaces (Detect.forward replacement, CerberusDetInference mirror) on a
stand-in model with the reference's head interface.  /root/reference is not available on the GPU box, so
the stand-in Detect class reproduces only the attributes the drop-in touches; its own forward is the
oracle's torch restatement of the reference eval branch."""
import copy

import pytest
import torch
import torch.nn as nn

from cerberusdet_b200.synth import STRIDES
from tol import check_decode

pytestmark = pytest.mark.gpu


class Detect(nn.Module):
    dynamic = False
    export = False
    shape = None
    anchors = torch.empty(0)
    strides = torch.empty(0)

    def __init__(self, nc, ch):
        super().__init__()
        self.nc, self.nl, self.reg_max = nc, len(ch), 16
        self.no = nc + 64
        self.stride = torch.tensor(STRIDES)
        self.cv2 = nn.ModuleList(nn.Conv2d(c, 64, 1) for c in ch)
        self.cv3 = nn.ModuleList(nn.Conv2d(c, nc, 1) for c in ch)

    def forward(self, x):  # torch restatement of reference models/yolo.py:87-100 (oracle code path)
        from oracle import ref_port as rp

        for i in range(self.nl):
            x[i] = torch.cat((self.cv2[i](x[i]), self.cv3[i](x[i])), 1)
        if self.training:
            return x
        y = rp.decode_port(x, self.nc, [float(s) for s in self.stride])
        return y if self.export else (y, x)


class TwoTaskModel(nn.Module):
    def __init__(self):
        super().__init__()
        self.stem = nn.Conv2d(3, 8, 3, padding=1)
        self.heads = nn.ModuleDict({"voc": Detect(20, (8, 8, 8)), "animals": Detect(19, (8, 8, 8))})
        self.names = {"voc": [f"v{i}" for i in range(20)], "animals": [f"a{i}" for i in range(19)]}
        self.stride = torch.tensor(STRIDES)
        for h in self.heads.values():  # spread the class logits so a few hundred candidates pass
            for m in h.cv3:
                nn.init.normal_(m.weight, std=1.5)
                nn.init.constant_(m.bias, -3.0)
            for m in h.cv2:
                nn.init.normal_(m.weight, std=1.0)

    def forward(self, x):
        f = self.stem(x)
        feats = [nn.functional.avg_pool2d(f, int(s)) for s in STRIDES]
        return {t: h([z.clone() for z in feats]) for t, h in self.heads.items()}


@pytest.fixture()
def patched():
    from cerberusdet_b200.detect import detect_forward

    Detect._cerb_reference_forward = Detect.forward
    orig = Detect.forward
    Detect.forward = detect_forward
    yield
    Detect.forward = orig
    del Detect._cerb_reference_forward


@pytest.mark.parametrize("half", [False, True])
def test_detect_forward_dropin(patched, half):
    torch.manual_seed(0)
    head = Detect(12, (8, 8, 8)).cuda().eval()
    if half:
        head.half()
    dt = torch.float16 if half else torch.float32
    feats = [torch.randn(2, 8, 32 // k, 48 // k, device="cuda", dtype=dt) for k in (1, 2, 4)]
    with torch.no_grad():
        y, x = head([f.clone() for f in feats])                       # B200 path
        y_ref, x_ref = Detect._cerb_reference_forward(head, [f.clone() for f in feats])  # torch path, same convs
    assert y.dtype == dt and tuple(y.shape) == (2, 16, 32 * 48 + 16 * 24 + 8 * 12)
    assert all(torch.equal(a, b) for a, b in zip(x, x_ref))           # the list holds the raw per-level tensors
    ok, msg = check_decode(y, y_ref.cpu() if False else y_ref, [t.shape[2:] for t in x], STRIDES, 12)
    assert ok, msg
    assert head.shape == feats[0].shape and tuple(head.anchors.shape) == (2, y.shape[2]) and tuple(head.strides.shape) == (1, y.shape[2])
    head.export = True
    with torch.no_grad():
        assert torch.equal(head([f.clone() for f in feats]), y)
    head.export = False
    head.train()
    out = head([f.clone() for f in feats])                            # training branch untouched: list of raw tensors
    assert isinstance(out, list) and len(out) == 3
    cpu_head = copy.deepcopy(head).float().cpu().eval()               # CPU tensors run the class's own forward
    with torch.no_grad():
        y_cpu, _ = cpu_head([f.float().cpu() for f in feats])
    assert not y_cpu.is_cuda


@pytest.mark.parametrize("half", [False, True])
def test_inference_mirror_matches_oracle_pipeline(patched, half):
    from cerberusdet_b200 import cross_task as ct
    from cerberusdet_b200.inference import CerberusDetInference
    from oracle import ref_port as rp

    torch.manual_seed(1)
    model = TwoTaskModel().cuda()
    eng = CerberusDetInference(model=model, device="cuda:0", conf_thres=0.25, iou_thres=0.45, half=half, img_size=64)
    assert eng.stride == 32 and len(eng.all_class_names) == 39 and eng.categories_inds_map["animals"][3] == 23
    x = torch.rand(3, 3, 96, 128, device="cuda")
    x = x.half() if half else x
    got = eng.predict(x, original_shape=[(480, 640), (96, 128), (300, 500)], max_det=50)
    assert len(got) == 3

    # oracle pipeline on the same conv outputs: decode on the GPU path is checked elsewhere; here selection
    # must be identical given the same decoded tensors, so feed the oracle NMS with the drop-in's own y
    with torch.no_grad():
        out = eng.model(x)
    per_task = {}
    for t, (y, _) in out.items():
        per_task[t] = rp.nms_port(y.cpu(), 0.25, 0.45, max_det=50, greedy="c")
    for i in range(3):
        det = ct.combine_tasks({t: per_task[t][i] for t in per_task}, eng.categories_inds_map)
        det = ct.suppress_between_tasks(det, eng.categories_inds_map, 0.8)
        if len(det):
            det[:, :4] = ct.rescale_boxes((96, 128), det[:, :4], [(480, 640), (96, 128), (300, 500)][i]).round()
        want = [{"box": [int(v) for v in r[:4]], "score": float(r[4]), "label": int(r[5])} for r in det.tolist()]
        assert len(got[i]) == len(want)
        for g, w in zip(got[i], want):
            assert g["box"] == w["box"] and g["label"] == w["label"] and g["score"] == w["score"]
            assert g["label_name"] == eng.all_class_names[g["label"]]
            assert g["task"] == ("voc" if g["label"] < 20 else "animals")
    assert sum(len(r) for r in got) > 0


def _random_task_dets(gen, T, B, max_det, clustered, ties):
    """Padded per-task NMS-like outputs with overlapping boxes across tasks (and exact score ties)."""
    dets = torch.zeros(T, B, max_det, 6)
    counts = torch.zeros(T, B, dtype=torch.int32)
    for b in range(B):
        k = 5
        centres = torch.rand(k, 2, generator=gen) * 400 + 100
        for t in range(T):
            n = int(torch.randint(0, max_det + 1, (1,), generator=gen))
            if b == 1 and t == 0:
                n = 0  # an empty task
            pick = torch.randint(0, k, (n,), generator=gen)
            c = centres[pick] + torch.randn(n, 2, generator=gen) * (3 if clustered else 90)
            wh = 50 + torch.rand(n, 2, generator=gen) * 30
            sc = torch.rand(n, generator=gen)
            if ties:
                sc = (sc * 8).round() / 8  # many exactly equal scores, also across tasks
            sc = sc.sort(descending=True).values
            dets[t, b, :n] = torch.cat((c - wh / 2, c + wh / 2, sc[:, None], torch.randint(0, 12, (n, 1), generator=gen).float()), 1)
            counts[t, b] = n
    return dets, counts


@pytest.mark.parametrize("clustered,ties", [(True, False), (True, True), (False, False)])
@pytest.mark.parametrize("thr", [0.8, 0.5, 0.05])
@pytest.mark.parametrize("with_scale", [False, True])
def test_cross_task_merge_matches_host_tail(clustered, ties, thr, with_scale):
    """GPU cross-task merge == cross_task.py (itself pinned to the reference's nms_between_tasks / scale_boxes
    by tests/test_host_logic.py), bit for bit, including equal-score ties and the all-deleted rule."""
    from cerberusdet_b200 import cross_task as ct
    from cerberusdet_b200.ops import cross_task_merge

    gen = torch.Generator().manual_seed(int(clustered) * 7 + int(ties) * 3 + int(thr * 100))
    T, B, md = 3, 5, 40
    dets, counts = _random_task_dets(gen, T, B, md, clustered, ties)
    names = {"a": ["x"] * 12, "b": ["y"] * 12, "c": ["z"] * 12}
    maps, _ = ct.category_maps(names)
    offsets = [0, 12, 24]
    shapes = [(480, 640), (1080, 1920), (333, 500), (640, 640), (100, 900)]
    scale = None
    if with_scale:
        rows = []
        for (oh, ow) in shapes:
            gain = min(640 / oh, 640 / ow)
            rows.append([gain, (640 - ow * gain) / 2, (640 - oh * gain) / 2, float(ow), float(oh)])
        scale = torch.tensor(rows, dtype=torch.float64).to(torch.float32)
    merged, mc = cross_task_merge(dets.cuda(), counts.cuda(), offsets, thr, scale)
    merged, mc = merged.cpu(), mc.cpu()
    for b in range(B):
        per_task = {t: dets[k, b, : int(counts[k, b])] for k, t in enumerate(names)}
        want = ct.combine_tasks(per_task, maps)
        want = ct.suppress_between_tasks(want, maps, thr)
        if with_scale and len(want):
            want[:, :4] = ct.rescale_boxes((640, 640), want[:, :4], shapes[b]).round()
        got = merged[b, : int(mc[b])]
        assert got.shape == want.shape, (b, got.shape, want.shape)
        assert torch.equal(got, want), b


@pytest.mark.parametrize("thr", [0.8, 0.3])
def test_cross_task_merge_beyond_1024_rows(thr):
    """max_det = 1000 (the reference's detect.py:124): T*max_det = 3000 rows per image.  Images whose rows exceed the
    shared-memory tables run out of the global workspace, the others out of shared memory -- same bits as the host tail."""
    from cerberusdet_b200 import cross_task as ct
    from cerberusdet_b200.ops import cross_task_merge

    gen = torch.Generator().manual_seed(11)
    T, B, md = 3, 3, 1000
    dets, counts = _random_task_dets(gen, T, B, md, clustered=False, ties=False)
    counts[:, 0] = torch.tensor([900, 700, 1000], dtype=torch.int32)   # 2600 rows: workspace path
    counts[:, 2] = torch.tensor([300, 0, 250], dtype=torch.int32)      # 550 rows: shared-memory path
    for b in (0, 2):
        for t in range(T):
            n = int(counts[t, b])
            c = torch.rand(n, 2, generator=gen) * 500 + 50
            wh = 30 + torch.rand(n, 2, generator=gen) * 60
            sc = torch.rand(n, generator=gen).sort(descending=True).values
            dets[t, b] = 0
            dets[t, b, :n] = torch.cat((c - wh / 2, c + wh / 2, sc[:, None], torch.randint(0, 12, (n, 1), generator=gen).float()), 1)
    names = {"a": ["x"] * 12, "b": ["y"] * 12, "c": ["z"] * 12}
    maps, _ = ct.category_maps(names)
    merged, mc = cross_task_merge(dets.cuda(), counts.cuda(), [0, 12, 24], thr)
    merged, mc = merged.cpu(), mc.cpu()
    for b in range(B):
        per_task = {t: dets[k, b, : int(counts[k, b])] for k, t in enumerate(names)}
        want = ct.suppress_between_tasks(ct.combine_tasks(per_task, maps), maps, thr)
        got = merged[b, : int(mc[b])]
        assert got.shape == want.shape, (b, got.shape, want.shape)
        assert torch.equal(got, want), b


def test_cross_task_merge_everything_deleted_keeps_everything():
    """Two tasks with one identical box each and equal scores: the column wins the tie and the row is deleted;
    with mutual total deletion the reference returns the input unchanged (general.py:551-552)."""
    from cerberusdet_b200 import cross_task as ct
    from cerberusdet_b200.ops import cross_task_merge

    dets = torch.zeros(2, 1, 4, 6)
    counts = torch.tensor([[1], [1]], dtype=torch.int32)
    dets[0, 0, 0] = torch.tensor([10, 10, 50, 50, 0.5, 1.0])
    dets[1, 0, 0] = torch.tensor([10, 10, 50, 50, 0.5, 2.0])
    maps, _ = ct.category_maps({"a": ["x"] * 3, "b": ["y"] * 3})
    merged, mc = cross_task_merge(dets.cuda(), counts.cuda(), [0, 3], 0.5)
    want = ct.suppress_between_tasks(ct.combine_tasks({"a": dets[0, 0, :1], "b": dets[1, 0, :1]}, maps), maps, 0.5)
    assert int(mc[0]) == want.shape[0] and torch.equal(merged[0, : int(mc[0])].cpu(), want)


def test_val_match_batch_matches_host_restatement():
    """GPU process_batch for a whole batch == val_stats.match_predictions per image (itself pinned to the reference's
    process_batch by tests/test_host_logic.py)."""
    from test_host_logic import _val_case

    from cerberusdet_b200.ops import match_batch
    from cerberusdet_b200.val_stats import match_predictions

    g = torch.Generator().manual_seed(3)
    iouv = torch.linspace(0.5, 0.95, 10)
    B, md = 9, 300
    dets = torch.zeros(B, md, 6)
    counts = torch.zeros(B, dtype=torch.int32)
    labs, offs, per = [], [0], []
    for b in range(B):
        M = 0 if b == 2 else int(torch.randint(1, 40, (1,), generator=g))
        N = 0 if b == 5 else (300 if b == 7 else int(torch.randint(1, 120, (1,), generator=g)))
        det, lab = _val_case(g, M, N, ncls=3)
        dets[b, :N] = det
        counts[b] = N
        labs.append(lab)
        offs.append(offs[-1] + M)
        per.append((det, lab))
    got = match_batch(dets.cuda(), counts.cuda(), torch.cat(labs, 0), offs, iouv).cpu()
    for b, (det, lab) in enumerate(per):
        want = match_predictions(det, lab, iouv)
        assert torch.equal(got[b, : det.shape[0]], want), b
        assert not got[b, det.shape[0]:].any()

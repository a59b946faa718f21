"""GPU parity of the head-tail fusion (``cerb_head_tail``: last 1x1 convolutions of both towers +
concat + eval decode in one tcgen05 kernel) against vectors the UNMODIFIED reference ``Detect`` module produced with its
real conv towers (tests/golden/headtail_*.npz, oracle/gen_golden_headtail.py), against the oracle port on seeded inputs,
and -- at bench shapes -- against the unfused path (cuDNN 1x1 convolutions + ``cerb_decode_split``).

Tolerance.  The fused kernel and the reference both compute ``half(fp32 sum + bias)`` for every conv output, but sum
in different orders, so a logit whose exact value lies within an fp32 rounding error of a half rounding boundary
can come out one half-ulp apart ("flip").  Everything downstream is the decode that the other tests pin.  So:
  * anchors all of whose outputs are bit-identical to the decode of the REFERENCE's raw conv outputs: >= 97 %;
  * the rest within what one-ulp logit flips can do: scores 1.6e-2 relative (a logit ulp is <= 2^-7 for |x| < 16 and
    d sigmoid / sigmoid <= dx) + 1e-6, boxes 0.06 grid units * stride (a DFL expectation moves by <= 15 * dp) on top of
    the decode tolerance of tests/tol.py."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden_manifest, golden_names, load_golden
from cerberusdet_b200.synth import STRIDES
from tol import box_atol_per_level

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from cerberusdet_b200 import _lib, ops as o

    _lib.load()
    return o


def check_head_tail(y, y_unfused, level_hw, strides, min_identical=0.97):
    """y: fused output; y_unfused: decode kernel on the reference's (or an exact) raw conv output.  (ok, message)."""
    yf, rf = y.float().cpu(), y_unfused.float().cpu()
    if not torch.isfinite(yf).all():
        return False, "non-finite output"
    same = (y.cpu() == y_unfused.cpu()).all(dim=1)  # [B, A]: anchors with every output bit-identical
    frac = same.float().mean().item()
    if frac < min_identical:
        per_image = [round(v, 3) for v in same.float().mean(dim=1).tolist()]
        return False, f"only {100 * frac:.2f} % of the anchors are bit-identical to the unfused path (per image: {per_image[:8]} ...)"
    off = 0
    for (h, w), s, atol in zip(level_hw, strides, box_atol_per_level(level_hw, strides, torch.float16)):
        n = h * w
        d = (yf[:, :4, off : off + n] - rf[:, :4, off : off + n]).abs()
        bound = 1e-3 * rf[:, :4, off : off + n].abs() + atol + 0.06 * s
        if (d > bound).any():
            return False, f"box mismatch at level offset {off}: worst excess {(d - bound).max().item():.3e}"
        off += n
    d = (yf[:, 4:] - rf[:, 4:]).abs()
    bound = 1.6e-2 * rf[:, 4:].abs() + 1e-6
    if (d > bound).any():
        return False, f"score mismatch: worst excess {(d - bound).max().item():.3e}"
    return True, f"{100 * frac:.2f} % of anchors bit-identical"


def _golden_case(name):
    g = load_golden(name)
    t = lambda k: torch.from_numpy(g[k]).cuda()  # noqa: E731
    return g, t


@pytest.mark.parametrize("name", [n for n in golden_names("headtail") if "f16" in n])
def test_head_tail_golden(ops, name):
    g, t = _golden_case(name)
    ref = torch.from_numpy(g["y"])
    nc = ref.shape[1] - 4
    feats = dict(box=[t(f"box_feat{l}") for l in range(3)], cls=[t(f"cls_feat{l}") for l in range(3)])
    y = ops.head_tail([feats["box"]], [feats["cls"]], [[t(f"box_w{l}") for l in range(3)]], [[t(f"box_b{l}") for l in range(3)]],
                      [[t(f"cls_w{l}") for l in range(3)]], [[t(f"cls_b{l}") for l in range(3)]], STRIDES)[0]
    torch.cuda.synchronize()
    assert y.dtype == ref.dtype and tuple(y.shape) == tuple(ref.shape)
    shapes = [tuple(x.shape[2:]) for x in feats["box"]]
    # the decode kernel on the reference's own raw conv outputs (itself pinned to g["y"] by test_gpu_parity)
    y_dec = ops.decode_heads([[t(f"raw{l}") for l in range(3)]], STRIDES)[0]
    ok, msg = check_head_tail(y, y_dec, shapes, STRIDES)
    assert ok, f"{name} vs decode(reference raw): {msg}"
    ok, msg = check_head_tail(y, ref.cuda(), shapes, STRIDES)
    assert ok, f"{name} vs reference y: {msg}"
    assert nc == golden_manifest()[name]["nc"]


def _random_case(B, ncs, c2, c3, imgsz, seed):
    gen = torch.Generator().manual_seed(seed)
    T = len(ncs)
    hw = [(imgsz[0] // int(s), imgsz[1] // int(s)) for s in STRIDES]
    mk = lambda *shape, std=1.0: (torch.randn(*shape, generator=gen) * std).half()  # noqa: E731
    case = dict(box=[], cls=[], bw=[], bb=[], cw=[], cb=[])
    for t in range(T):
        case["box"].append([mk(B, c2, h, w) for h, w in hw])
        case["cls"].append([mk(B, c3, h, w) for h, w in hw])
        case["bw"].append([mk(64, c2, 1, 1, std=3.0 / c2**0.5) for _ in hw])
        case["bb"].append([mk(64, std=1.0) for _ in hw])
        case["cw"].append([mk(ncs[t], c3, 1, 1, std=2.0 / c3**0.5) for _ in hw])
        case["cb"].append([(mk(ncs[t], std=1.0) - 5.0).half() for _ in hw])
    return case, hw


def _cuda(case):
    return {k: [[x.cuda() for x in lv] for lv in v] for k, v in case.items()}


@pytest.mark.parametrize("B,ncs,c2,c3,imgsz", [
    (2, [20, 19, 12], 80, 320, (640, 640)),    # yolov8x widths, the bench's task heads
    (3, [12], 64, 64, (256, 192)),              # yolov8n widths, one task
    (1, [80, 33], 64, 128, (320, 320)),         # nc = 80: six class chunks; P5 is 10x10 -> H*W % 8 != 0 is rejected below
    (2, [20, 1], 64, 192, (512, 384)),          # nc = 1
])
def test_head_tail_vs_oracle_port(ops, B, ncs, c2, c3, imgsz):
    """Seeded inputs: the fused kernel against oracle/ref_port.head_tail_port (torch CPU convolutions + the decode port)."""
    from oracle import ref_port as rp

    case, hw = _random_case(B, ncs, c2, c3, imgsz, seed=sum(ncs) + c3)
    d = _cuda(case)
    if any((h * w) % 8 for h, w in hw):
        with pytest.raises(ValueError, match="multiple of 8"):
            ops.head_tail(d["box"], d["cls"], d["bw"], d["bb"], d["cw"], d["cb"], STRIDES)
        return
    ys = ops.head_tail(d["box"], d["cls"], d["bw"], d["bb"], d["cw"], d["cb"], STRIDES)
    torch.cuda.synchronize()
    for t in range(len(ncs)):
        _, raw = rp.head_tail_port(case["box"][t], case["cls"][t], case["bw"][t], case["bb"][t], case["cw"][t], case["cb"][t], STRIDES)
        y_dec = ops.decode_heads([[x.cuda() for x in raw]], STRIDES)[0]
        ok, msg = check_head_tail(ys[t], y_dec, hw, STRIDES)
        assert ok, f"task {t}: {msg}"


def test_head_tail_bench_shape_vs_unfused_and_nms(ops):
    """BASELINE config 3 shape (B=64, 3 task heads, 640x640, yolov8x widths): fused == (exact convolution in fp64,
    rounded once to half) + cerb_decode_split within the flip tolerance, and == cuDNN's half convolutions + cerb_decode_split
    in value (cuDNN's own rounding flips are not ours to bound); the score summary it writes is exactly the maximum of
    every 8-anchor score vector of ITS y, and NMS on the fused output with / without that summary selects the same rows."""
    B, ncs, c2, c3 = 64, [20, 19, 12], 80, 320
    case, hw = _random_case(B, ncs, c2, c3, (640, 640), seed=7)
    d = _cuda(case)
    del case
    ys = ops.head_tail(d["box"], d["cls"], d["bw"], d["bb"], d["cw"], d["cb"], STRIDES)

    def exact(x, w, b):  # [B, c, H, W] x [n, c, 1, 1]: fp64 sum, one rounding to half
        out = torch.empty((x.shape[0], w.shape[0]) + tuple(x.shape[2:]), dtype=torch.float16, device=x.device)
        w64, b64 = w.reshape(w.shape[0], -1).double(), b.double()
        for i in range(0, x.shape[0], 8):
            out[i : i + 8] = (torch.einsum("nc,bchw->bnhw", w64, x[i : i + 8].double()) + b64[None, :, None, None]).half()
        return out

    for t in range(3):
        box = [exact(d["box"][t][l], d["bw"][t][l], d["bb"][t][l]) for l in range(3)]
        cls = [exact(d["cls"][t][l], d["cw"][t][l], d["cb"][t][l]) for l in range(3)]
        ye = ops.decode_heads_split([box], [cls], STRIDES)[0]
        ok, msg = check_head_tail(ys[t], ye, hw, STRIDES)
        assert ok, f"task {t} vs exact convolution: {msg}"
        box = [F.conv2d(d["box"][t][l], d["bw"][t][l], d["bb"][t][l]) for l in range(3)]
        cls = [F.conv2d(d["cls"][t][l], d["cw"][t][l], d["cb"][t][l]) for l in range(3)]
        yu = ops.decode_heads_split([box], [cls], STRIDES)[0]
        ok, msg = check_head_tail(ys[t], yu, hw, STRIDES, min_identical=0.0)
        assert ok, f"task {t} vs cuDNN convolution: {msg}"
        print(f"task {t}: anchors bit-identical to cuDNN + decode: {100 * (ys[t] == yu).all(dim=1).float().mean().item():.2f} %, "
              f"to exact + decode: {100 * (ys[t] == ye).all(dim=1).float().mean().item():.2f} %")
        del box, cls, ye, yu
    torch.cuda.synchronize()
    for t in range(3):
        sm = ops.find_summary(ys[t])
        assert sm is not None
        want = ys[t][:, 4:].reshape(B, ncs[t], -1, 8).amax(dim=-1)
        assert torch.equal(sm[..., : want.shape[-1]], want), f"task {t}: score summary is not the 8-anchor maximum of y"
    kw = dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)
    d1, c1 = ops.nms_batched(ys, **kw)
    d2, c2_ = ops.nms_batched(ys, use_summary=False, **kw)
    assert torch.equal(c1, c2_) and torch.equal(d1, d2)
    assert int(c1.min()) > 0


def test_head_tail_rejects_what_it_cannot_run(ops):
    case, _ = _random_case(1, [4], 64, 64, (128, 128), seed=1)
    d = _cuda(case)
    f32 = {k: [[x.float() for x in lv] for lv in v] for k, v in d.items()}
    with pytest.raises(TypeError, match="float16"):
        ops.head_tail(f32["box"], f32["cls"], f32["bw"], f32["bb"], f32["cw"], f32["cb"], STRIDES)
    bad = [[x[:, :40].contiguous() for x in lv] for lv in d["box"]]  # c2 = 40: not a multiple of 16
    with pytest.raises(ValueError):
        ops.head_tail(bad, d["cls"], [[w[:, :40].contiguous() for w in lv] for lv in d["bw"]], d["bb"], d["cw"], d["cb"], STRIDES)
    with pytest.raises(TypeError, match="CUDA"):
        ops.head_tail([[x.cpu() for x in lv] for lv in d["box"]], d["cls"], d["bw"], d["bb"], d["cw"], d["cb"], STRIDES)

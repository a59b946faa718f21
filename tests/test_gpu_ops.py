"""GPU tests of the torch custom ops (north_star: "a thin C-ABI shim registered as a torch custom op") and of the streaming
engine built on them (cerberusdet_b200/pipeline.py)."""
import pytest
import torch

from cerberusdet_b200.synth import STRIDES, synth_heads

pytestmark = pytest.mark.gpu

KW = dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)


def _heads(bsz=2, ncs=(20, 19, 12), imgsz=256, dtype=torch.float16, cfg=61):
    return [[x.cuda() for x in lv] for lv in synth_heads(range(bsz), list(ncs), imgsz, dtype, "iid", cfg=cfg)]


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_opcheck_all_registered_ops(dtype):
    """torch.library.opcheck: schema, fake-tensor kernel (shapes / dtypes / devices / no aliasing) and the dispatch
    registrations of every cerb:: op, on real arguments."""
    from cerberusdet_b200 import ops

    heads = _heads(dtype=dtype)
    flat = [x for lv in heads for x in lv]
    nc, strides = [20, 19, 12], [float(s) for s in STRIDES]
    tests = ("test_schema", "test_faketensor", "test_autograd_registration")
    torch.library.opcheck(torch.ops.cerb.decode.default, (flat, nc, strides), test_utils=tests)
    box = [x[:, :64].contiguous() for x in flat]
    cls = [x[:, 64:].contiguous() for x in flat]
    torch.library.opcheck(torch.ops.cerb.decode_split.default, (box, cls, nc, strides), test_utils=tests)
    out = ops.decode_op(flat, nc, strides)
    ys, sms = out[:3], out[3:]
    assert all(s.shape[-1] > 0 for s in sms)  # 256x256: every level (1024 / 256 / 64 anchors) allows 16-byte vectors
    nms_args = (ys, 0.001, 0.6, None, False, True, 300, 30000, 7680.0, sms)
    torch.library.opcheck(torch.ops.cerb.nms.default, nms_args, test_utils=tests)
    torch.library.opcheck(torch.ops.cerb.nms.default, (ys, 0.25, 0.45, [0, 3], True, False, 50, 30000, 7680.0, []), test_utils=tests)
    dets = torch.empty((3, 2, 300, 6), device="cuda")
    counts = torch.empty((3, 2), dtype=torch.int32, device="cuda")
    torch.library.opcheck(torch.ops.cerb.nms_out.default, nms_args + (dets, counts), test_utils=tests)
    torch.library.opcheck(torch.ops.cerb.decode_out.default, (flat, nc, strides, ops.decode_buffers(heads)), test_utils=tests)
    torch.library.opcheck(torch.ops.cerb.decode_nms.default, (flat, nc, strides, 0.001, 0.6, None, False, True, 300, 30000, 7680.0),
                          test_utils=tests)
    # shapes that rule the score summary out: the fake kernel and the real one must still agree
    odd = [[x.cuda() for x in lv] for lv in synth_heads(range(2), [5], (40, 24), dtype, "iid", cfg=62, strides=(8.0,))]
    torch.library.opcheck(torch.ops.cerb.decode.default, ([odd[0][0]], [5], [8.0]), test_utils=tests)
    o = ops.decode_op([odd[0][0]], [5], [8.0])
    assert o[1].shape == (2, 5, 0)


def test_ops_under_torch_compile_match_eager():
    """One compiled function over the registered ops (aot_eager: no code generation, the ops stay the hot path)."""
    from cerberusdet_b200 import ops

    heads = _heads()
    flat = [x for lv in heads for x in lv]
    nc, strides = [20, 19, 12], [float(s) for s in STRIDES]

    def fn(levels):
        out = torch.ops.cerb.decode(levels, nc, strides)
        dets, counts = torch.ops.cerb.nms(out[:3], 0.001, 0.6, None, False, True, 300, 30000, 7680.0, out[3:])
        return dets, counts, torch.ops.cerb.decode_nms(levels, nc, strides, 0.001, 0.6, None, False, True, 300, 30000, 7680.0)

    d0, c0, fused0 = fn(flat)
    d1, c1, fused1 = torch.compile(fn, backend="aot_eager", fullgraph=True)(flat)
    assert torch.equal(d0, d1) and torch.equal(c0, c1)
    assert torch.equal(fused0[0], d0) and torch.equal(fused0[1], c0) and torch.equal(fused1[0], d0) and torch.equal(fused1[1], c0)
    ys = ops.decode_heads(heads, STRIDES)
    for t in range(3):
        assert torch.equal(fused1[2 + t], ys[t])


def test_decode_out_buffers_and_inference_mode():
    from cerberusdet_b200 import ops

    heads = _heads()
    ys = ops.decode_heads(heads, STRIDES)
    buf = ops.decode_buffers(heads)
    ys2 = ops.decode_heads(heads, STRIDES, out=buf)
    assert all(a.data_ptr() == b.data_ptr() for a, b in zip(ys2, buf[:3]))
    assert all(torch.equal(a, b) for a, b in zip(ys, ys2))
    assert all(ops.find_summary(y) is not None for y in ys2)
    want = ops.nms_batched(ys, **KW)
    got = ops.nms_batched(ys2, **KW)
    assert torch.equal(want[0], got[0]) and torch.equal(want[1], got[1])
    # an in-place write invalidates the attached summary (the version counter moved)
    ys2[0].mul_(1.0)
    assert ops.find_summary(ys2[0]) is None
    # tensors born under inference_mode carry no version counter: no summary, same results
    with torch.inference_mode():
        yi = ops.decode_heads(heads, STRIDES)
        assert all(ops.find_summary(y) is None for y in yi)
        gi = ops.nms_batched(yi, **KW)
    assert torch.equal(gi[0], want[0]) and torch.equal(gi[1], want[1])
    with pytest.raises(ValueError):
        ops.decode_heads(heads, STRIDES, out=buf[:-1])


@pytest.mark.parametrize("overlap,schedule", [(True, "pdl"), (True, "streams"), (False, "pdl")])
def test_pipeline_equals_direct_calls(overlap, schedule, monkeypatch):
    """The streaming engine (static buffers, CUDA graphs; decode of batch k beside the NMS of batch k-1 -- on one stream
    with a programmatic launch, or on two graph branches) gives the bits of the plain two-call path, batch after batch,
    also when the static inputs change between steps."""
    from cerberusdet_b200 import ops, pipeline
    from cerberusdet_b200.pipeline import PostHeadPipeline

    monkeypatch.setattr(pipeline, "_SCHEDULE", schedule)

    batches = [_heads(bsz=4, cfg=70 + i) for i in range(4)]
    static = [[x.clone() for x in lv] for lv in batches[0]]
    pipe = PostHeadPipeline(static, STRIDES, KW, timed_parities=[1, 0], overlap=overlap)
    want = []
    for hb in batches:
        ys = ops.decode_heads(hb, STRIDES)
        want.append(ops.nms_batched(ys, **KW))
    got = {}
    for k, hb in enumerate(batches):
        for s, x in zip([t for lv in static for t in lv], [t for lv in hb for t in lv]):
            s.copy_(x)
        done = pipe.step(timed=(0 if k == 1 else 1 if k == 2 else None) if overlap else None)
        if done is not None:
            torch.cuda.synchronize()
            batch = k if not overlap else k - 1
            got[batch] = (pipe.outs[done][0].clone(), pipe.outs[done][1].clone())
    done = pipe.flush()
    if done is not None:
        torch.cuda.synchronize()
        got[len(batches) - 1] = (pipe.outs[done][0].clone(), pipe.outs[done][1].clone())
    assert sorted(got) == list(range(len(batches)))
    for k in range(len(batches)):
        assert torch.equal(got[k][1], want[k][1]) and torch.equal(got[k][0], want[k][0]), f"batch {k}"
    if overlap:
        for i in range(2):
            nms_ms, dec_ms = pipe.timed_ms(i)
            assert 0 < nms_ms < 50 and 0 < dec_ms < 50


def test_nms_statistics_counters():
    from cerberusdet_b200 import ops

    heads = _heads(bsz=3)
    ys = ops.decode_heads(heads, STRIDES)
    d0, c0 = ops.nms_batched(ys, **KW)
    d1, c1, stats = ops.nms_statistics(ys, **KW)
    assert torch.equal(d0, d1) and torch.equal(c0, c1)
    s = stats.cpu()
    assert tuple(s.shape) == (3, 3, 2)
    assert (s[..., 1] >= c0.cpu()).all() and (s[..., 1] <= 30000).all()  # consumed >= kept
    assert (s[..., 0] > 0).all()

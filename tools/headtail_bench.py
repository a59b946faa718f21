"""Head-tail fusion (cerb_head_tail) against the unfused sequence on the same GPU, BASELINE config 3 shape with yolov8x
head widths (c2 = 80, c3 = 320): CUDA-graph replays, events around the replay.

    python tools/headtail_bench.py [--batch 64] [--reps 20]

unfused A = what the reference runs:   18 cuDNN 1x1 convolutions + 9 torch.cat + cerb_decode (raw heads written, read twice)
unfused B = this repo without fusion:  18 cuDNN 1x1 convolutions + cerb_decode_split (no cat)
fused     = cerb_head_tail (one persistent tcgen05 kernel; the raw heads never exist)
Engineering numbers for profiles/r02_head_tail.md, not bench values."""
import argparse
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from cerberusdet_b200 import ops  # noqa: E402
from cerberusdet_b200.synth import STRIDES  # noqa: E402


def timed_graph(fn, reps):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s), torch.cuda.graph(g, stream=s):
        keep = fn()
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(3):
        g.replay()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    del keep
    return round(statistics.median(ts), 1), round(min(ts), 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--c2", type=int, default=80)
    ap.add_argument("--c3", type=int, default=320)
    args = ap.parse_args()
    B, ncs, c2, c3 = args.batch, [20, 19, 12], args.c2, args.c3
    hw = [(80, 80), (40, 40), (20, 20)]
    A = sum(h * w for h, w in hw)
    gen = torch.Generator(device="cuda").manual_seed(3)
    mk = lambda *shape, std=1.0: (torch.randn(*shape, generator=gen, device="cuda") * std).half()  # noqa: E731
    box = [[mk(B, c2, h, w) for h, w in hw] for _ in ncs]
    cls = [[mk(B, c3, h, w) for h, w in hw] for _ in ncs]
    bw = [[mk(64, c2, 1, 1, std=3.0 / c2**0.5) for _ in hw] for _ in ncs]
    bb = [[mk(64) for _ in hw] for _ in ncs]
    cw = [[mk(n, c3, 1, 1, std=2.0 / c3**0.5) for _ in hw] for n in ncs]
    cb = [[mk(n) - 5.0 for _ in hw] for n in ncs]

    def convs():
        return ([[F.conv2d(box[t][l], bw[t][l], bb[t][l]) for l in range(3)] for t in range(3)],
                [[F.conv2d(cls[t][l], cw[t][l], cb[t][l]) for l in range(3)] for t in range(3)])

    def unfused_a():
        bo, cl = convs()
        raw = [[torch.cat((bo[t][l], cl[t][l]), 1) for l in range(3)] for t in range(3)]
        return ops.decode_heads(raw, STRIDES)

    def unfused_b():
        bo, cl = convs()
        return ops.decode_heads_split(bo, cl, STRIDES)

    def fused():
        return ops.head_tail(box, cls, bw, bb, cw, cb, STRIDES)

    in_bytes = B * A * 2 * (c2 + c3) * len(ncs)
    out_bytes = sum(B * A * 2 * (4 + n) + B * n * (A // 8) * 2 for n in ncs)
    raw_bytes = sum(B * A * 2 * (64 + n) for n in ncs)
    res = {"batch": B, "c2": c2, "c3": c3, "input_MB": in_bytes / 1e6, "output_MB": out_bytes / 1e6, "raw_head_MB": raw_bytes / 1e6}
    for name, fn in (("convs_only", convs), ("unfused_reference_order", unfused_a), ("unfused_split", unfused_b), ("fused", fused)):
        med, mn = timed_graph(fn, args.reps)
        res[name + "_us"] = med
        res[name + "_us_min"] = mn
    res["fused_GBps"] = round((in_bytes + out_bytes) / (res["fused_us"] * 1e-6) / 1e9, 1)
    res["speedup_vs_reference_order"] = round(res["unfused_reference_order_us"] / res["fused_us"], 3)
    res["speedup_vs_split"] = round(res["unfused_split_us"] / res["fused_us"], 3)
    print(json.dumps(res))


if __name__ == "__main__":
    main()

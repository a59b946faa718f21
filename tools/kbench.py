"""Kernel-only timings by CUDA-graph replay (no host launch gaps): decode alone, NMS alone, decode -> NMS serial, and
decode(k+1) running concurrently with NMS(k) on a second stream.

    python tools/kbench.py [cfg3 cfg2 cfg4 cfg3f32 cfg3planted ...] [--chain 10] [--reps 8]
    CERB_LIB=cerberusdet_b200/libcerb_x.so python tools/kbench.py cfg3      # a variant build (tools/decode_variants.py style)

Every graph holds `chain` back-to-back launches of the thing measured; the inputs (261 MB at config 3) exceed the L2, so
consecutive launches flush each other.  Prints one JSON line per config.  Numbers from here are engineering numbers
for profiles/*.md, not bench values.
"""
import argparse
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cerberusdet_b200 import ops  # noqa: E402
from cerberusdet_b200.synth import STRIDES, synth_heads  # noqa: E402

VAL = dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)
CFG = {
    "cfg2": dict(ncs=[20, 19], B=32, imgsz=640, dtype=torch.float16, kw=dict(conf_thres=0.3, iou_thres=0.45, max_det=300)),
    "cfg3": dict(ncs=[20, 19, 12], B=64, imgsz=640, dtype=torch.float16, kw=VAL),
    "cfg3f32": dict(ncs=[20, 19, 12], B=64, imgsz=640, dtype=torch.float32, kw=VAL),
    "cfg4": dict(ncs=[20, 19, 12], B=16, imgsz=1280, dtype=torch.float16, kw=VAL),
    "cfg3planted": dict(ncs=[20, 19, 12], B=16, imgsz=640, dtype=torch.float16, regime="planted", kw=VAL),
    "cfg3planted64": dict(ncs=[20, 19, 12], B=64, imgsz=640, dtype=torch.float16, regime="planted", kw=VAL),
}


def graph_of(fn, chain, stream):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        keep = [fn() for _ in range(chain)]
    return g, keep


def time_graph(g, chain, reps):
    for _ in range(2):
        g.replay()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / chain)
    return round(statistics.median(ts), 2), round(min(ts), 2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cfgs", nargs="*", default=["cfg3"])
    ap.add_argument("--chain", type=int, default=10)
    ap.add_argument("--reps", type=int, default=8)
    ap.add_argument("--no-overlap", action="store_true")
    args = ap.parse_args()
    side = torch.cuda.Stream()
    side2 = torch.cuda.Stream()
    for name in args.cfgs:
        c = CFG[name]
        heads = synth_heads(range(c["B"]), c["ncs"], c["imgsz"], c["dtype"], c.get("regime", "iid"), cfg=3)
        dev = [[x.cuda() for x in lv] for lv in heads]
        A = sum(x.shape[2] * x.shape[3] for x in dev[0])
        elt = dev[0][0].element_size()
        byts = c["B"] * sum((64 + nc + 4 + nc) * elt * A for nc in c["ncs"])
        ys = ops.decode_heads(dev, STRIDES)
        dets, counts = ops.nms_batched(ys, **c["kw"])
        torch.cuda.synchronize()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            g_dec, _k1 = graph_of(lambda: ops.decode_heads(dev, STRIDES), args.chain, side)
            g_nms, _k2 = graph_of(lambda: ops.nms_batched(ys, **c["kw"]), args.chain, side)

            def both():
                y = ops.decode_heads(dev, STRIDES)
                return y, ops.nms_batched(y, **c["kw"])

            g_both, _k3 = graph_of(both, args.chain, side)
            g_ovl = None
            if not args.no_overlap:
                # software pipeline: step k decodes batch k+1 while the NMS of batch k runs on a second stream
                g_ovl = torch.cuda.CUDAGraph()
                keep = []
                with torch.cuda.graph(g_ovl, stream=side):
                    yprev = ys
                    for _ in range(args.chain):
                        side2.wait_stream(side)
                        with torch.cuda.stream(side2):
                            keep.append(ops.nms_batched(yprev, **c["kw"]))
                        ynew = ops.decode_heads(dev, STRIDES)
                        side.wait_stream(side2)
                        keep.append(ynew)
                        yprev = ynew
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        d = time_graph(g_dec, args.chain, args.reps)
        n = time_graph(g_nms, args.chain, args.reps)
        bth = time_graph(g_both, args.chain, args.reps)
        out = {"cfg": name, "decode_us": d[0], "decode_us_min": d[1], "decode_GBps": round(byts / d[0] / 1e3, 1),
               "decode_frac_8000": round(byts / d[0] / 1e3 / 8000, 3), "nms_us": n[0], "nms_us_min": n[1],
               "serial_us": bth[0], "serial_us_min": bth[1], "mean_kept": float(counts.float().mean())}
        if g_ovl is not None:
            o = time_graph(g_ovl, args.chain, args.reps)
            out.update({"overlap_us": o[0], "overlap_us_min": o[1]})
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()

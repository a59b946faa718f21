"""Per-kernel timings on the GPU box for the BASELINE configs (CUDA events, L2-sized inputs).
Usage: python tools/microbench.py [cfg2 cfg3 cfg4 ...]"""
import os, sys, statistics, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cerberusdet_b200 import ops
from cerberusdet_b200.synth import STRIDES, synth_heads

CFG = {
    "cfg2": dict(ncs=[20, 19], B=32, imgsz=640, dtype=torch.float16, kw=dict(conf_thres=0.3, iou_thres=0.45, max_det=300)),
    "cfg3": dict(ncs=[20, 19, 12], B=64, imgsz=640, dtype=torch.float16, kw=dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)),
    "cfg3f32": dict(ncs=[20, 19, 12], B=64, imgsz=640, dtype=torch.float32, kw=dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)),
    "cfg4": dict(ncs=[20, 19, 12], B=16, imgsz=1280, dtype=torch.float16, kw=dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)),
    "cfg3planted": dict(ncs=[20, 19, 12], B=16, imgsz=640, dtype=torch.float16, regime="planted", kw=dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)),
}

def timeit(fn, n=30, warm=5, flush=None):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        if flush is not None: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts), min(ts)

def main():
    names = sys.argv[1:] or ["cfg2", "cfg3", "cfg4"]
    if os.environ.get("CERB_CHUNKING"):  # "cap,first" -> test hook, e.g. CERB_CHUNKING=2048,369
        from cerberusdet_b200 import _lib
        cap, first = (int(v) for v in os.environ["CERB_CHUNKING"].split(","))
        assert _lib.load().cerb_debug_set_chunking(cap, first) == 0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name in names:
        c = CFG[name]
        heads = synth_heads(range(c["B"]), c["ncs"], c["imgsz"], c["dtype"], c.get("regime", "iid"), cfg=3)
        dev = [[x.cuda() for x in lv] for lv in heads]
        A = sum(x.shape[2] * x.shape[3] for x in dev[0])
        elt = dev[0][0].element_size()
        byts = c["B"] * sum((64 + nc + 4 + nc) * elt * A for nc in c["ncs"])
        ys = ops.decode_heads(dev, STRIDES)
        d_med, d_min = timeit(lambda: ops.decode_heads(dev, STRIDES), flush=flush)
        n_med, n_min = timeit(lambda: ops.nms_batched(ys, **c["kw"]), flush=flush)
        dets, counts = ops.nms_batched(ys, **c["kw"])
        print(json.dumps({"cfg": name, "decode_us_med": round(d_med, 1), "decode_us_min": round(d_min, 1),
                          "decode_GBps_med": round(byts / d_med / 1e3, 1), "nms_us_med": round(n_med, 1),
                          "nms_us_min": round(n_min, 1), "mean_kept": float(counts.float().mean())}), flush=True)

if __name__ == "__main__":
    main()

"""Training-time sibling decode (Loss.bbox_decode forward / backward) on the GPU box: CUDA events, L2 flushed between
iterations; algorithmic bytes = (64 + 4) * s forward, (64 + 4 + 64) * s backward per anchor.  Also times the same
ops in eager PyTorch on the same GPU (the reference's own code path on CUDA) for scale.
Usage: python tools/microbench_train.py"""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cerberusdet_b200 import ops


def timeit(fn, flush, n=30, warm=5):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)


def eager(ap, pred):  # reference utils/loss.py:126-131 + utils/tal.py:196-205, verbatim ops
    b, a, c = pred.shape
    proj = torch.arange(16, dtype=torch.float, device=pred.device)
    d = pred.view(b, a, 4, c // 4).softmax(3).matmul(proj.type(pred.dtype))
    lt, rb = torch.split(d, 2, -1)
    return torch.cat((ap - lt, ap + rb), -1)


def main():
    peak = 6531.9
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    B, A = 64, 8400
    for dtype in (torch.float16, torch.float32):
        s = 2 if dtype == torch.float16 else 4
        pred = torch.randn(B, A, 64, device="cuda").mul_(3).to(dtype).requires_grad_(True)
        ap = torch.rand(A, 2, device="cuda").mul_(80).to(dtype)
        go = torch.randn(B, A, 4, device="cuda").to(dtype)
        out = ops.bbox_decode(ap, pred)
        f = timeit(lambda: ops.bbox_decode(ap, pred.detach()), flush)
        bw = timeit(lambda: torch.autograd.grad(out, pred, go, retain_graph=True), flush)
        out_e = eager(ap, pred)
        fe = timeit(lambda: eager(ap, pred.detach()), flush)
        be = timeit(lambda: torch.autograd.grad(out_e, pred, go, retain_graph=True), flush)
        fb, bb = B * A * 68 * s, B * A * 132 * s
        print(json.dumps({"dtype": str(dtype), "B": B, "A": A, "fwd_us": round(f, 1), "fwd_GBps": round(fb / f / 1e3, 1),
                          "fwd_frac_of_measured": round(fb / f / 1e3 / peak, 3), "bwd_us": round(bw, 1),
                          "bwd_GBps": round(bb / bw / 1e3, 1), "bwd_frac_of_measured": round(bb / bw / 1e3 / peak, 3),
                          "eager_fwd_us": round(fe, 1), "eager_bwd_us": round(be, 1)}), flush=True)


if __name__ == "__main__":
    main()

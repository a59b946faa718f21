"""Per-CTA wall times of nms_kernel (one CTA = one (task, image) segment) with the -DNMS_BLOCK_TIMES build.
Build: CERB_OUT=cerberusdet_b200/libcerb_bt.so sh cerberusdet_b200/csrc/build.sh -DNMS_BLOCK_TIMES
Run:   CERB_LIB=cerberusdet_b200/libcerb_bt.so python tools/nms_block_times.py cfg2 cfg3 cfg4"""
import ctypes, os, statistics, sys
from collections import Counter
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cerberusdet_b200 import _lib, ops
from cerberusdet_b200.synth import STRIDES, synth_heads
from microbench import CFG

lib = _lib.load()
for name in sys.argv[1:] or ["cfg3"]:
    c = CFG[name]
    heads = synth_heads(range(c["B"]), c["ncs"], c["imgsz"], c["dtype"], c.get("regime", "iid"), cfg=3)
    dev = [[x.cuda() for x in lv] for lv in heads]
    n = c["B"] * len(c["ncs"])
    buf = (ctypes.c_ulonglong * (3 * n))()
    for rep in range(4):  # last repetition is the one read; decode right before, as in the bench
        ys = ops.decode_heads(dev, STRIDES)
        ops.nms_batched(ys, **c["kw"])
    assert lib.cerb_debug_read_block_times(buf, n) == 0
    st, en, sm = list(buf[0:n]), list(buf[n:2 * n]), list(buf[2 * n:3 * n])
    t0 = min(st)
    dur = [(e - s) / 1e3 for s, e in zip(st, en)]
    off = [(s - t0) / 1e3 for s in st]
    per_sm = Counter(sm)
    shared = [d for d, m in zip(dur, sm) if per_sm[m] > 1]
    alone = [d for d, m in zip(dur, sm) if per_sm[m] == 1]
    q = lambda v, p: sorted(v)[min(len(v) - 1, int(p * len(v)))]
    print(f"{name}: {n} CTAs on {len(per_sm)} SMs; kernel span {(max(en) - t0) / 1e3:.1f} us; start offsets max {max(off):.1f} us")
    print(f"   CTA duration us: min {min(dur):.1f}  p10 {q(dur, .1):.1f}  median {statistics.median(dur):.1f}  p90 {q(dur, .9):.1f}  max {max(dur):.1f}")
    if shared and alone:
        print(f"   alone on its SM ({len(alone)}): median {statistics.median(alone):.1f} max {max(alone):.1f};  sharing an SM ({len(shared)}): median {statistics.median(shared):.1f} max {max(shared):.1f}")
    worst = sorted(range(n), key=lambda i: -dur[i])[:5]
    print("   slowest segments (task, image, us):", [(i // c["B"], i % c["B"], round(dur[i], 1)) for i in worst])

"""Digest of the decode outputs (y and the score summary) on seeded inputs, to prove that two builds of the library
(CERB_LIB=...) or two kernels (CERB_DEBUG_DECODE_PIPE=0/1/2/4) are bit-identical.
Usage: [CERB_LIB=cerberusdet_b200/libX.so] python tools/decode_hash.py"""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cerberusdet_b200 import ops
from cerberusdet_b200.synth import STRIDES, synth_heads

for dtype in (torch.float16, torch.float32):
    heads = synth_heads(range(8), [20, 19, 12], 640, dtype, "iid", cfg=3)
    ys = ops.decode_heads([[x.cuda() for x in lv] for lv in heads], STRIDES)
    h = hashlib.sha256()
    for y in ys:
        h.update(y.cpu().numpy().tobytes())
        sm = ops.find_summary(y)
        if sm is not None:
            h.update(sm[..., : y.shape[2] // (8 if dtype == torch.float16 else 4)].contiguous().cpu().numpy().tobytes())
    print(str(dtype), h.hexdigest()[:16], flush=True)

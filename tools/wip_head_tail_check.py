"""First GPU check of the round-2 DRAFT kernel (cerberusdet_b200/csrc/wip/head_tail.cu) against the golden vectors the
unmodified reference Detect module produced (tests/golden/headtail_f16_*.npz).  Not part of the test suite.
    sh cerberusdet_b200/csrc/wip/build_wip.sh && python tools/wip_head_tail_check.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from tol import check_decode

lib = ctypes.CDLL(os.path.join(ROOT, "cerberusdet_b200", "libcerb_wip.so"))
vp, i, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
lib.cerb_wip_head_tail.argtypes = [vp] * 6 + [i] * 6 + [f, i, i, vp, vp, vp]
lib.cerb_wip_head_tail.restype = i
STRIDES = (8.0, 16.0, 32.0)
for name in ("headtail_f16_nc12", "headtail_f16_nc20_v8x"):
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))
    t = lambda k: torch.from_numpy(g[k]).cuda().contiguous()  # noqa: E731
    ref = torch.from_numpy(g["y"])
    B, no4, A = ref.shape
    nc = no4 - 4
    y = torch.zeros((B, no4, A), dtype=torch.float16, device="cuda")
    aoff, shapes = 0, []
    for l in range(3):
        u2, u3 = t(f"box_feat{l}"), t(f"cls_feat{l}")
        w2, w3 = t(f"box_w{l}").reshape(64, -1).contiguous(), t(f"cls_w{l}").reshape(nc, -1).contiguous()
        b2, b3 = t(f"box_b{l}"), t(f"cls_b{l}")
        H, W = u2.shape[2], u2.shape[3]
        rc = lib.cerb_wip_head_tail(u2.data_ptr(), u3.data_ptr(), w2.data_ptr(), b2.data_ptr(), w3.data_ptr(), b3.data_ptr(),
                                    B, u2.shape[1], u3.shape[1], nc, H, W, STRIDES[l], A, aoff, y.data_ptr(), None,
                                    torch.cuda.current_stream().cuda_stream)
        assert rc == 0, f"{name} level {l}: rc {rc}"
        torch.cuda.synchronize()
        shapes.append((H, W))
        aoff += H * W
    ok, msg = check_decode(y, ref, shapes, STRIDES, nc)
    nd = (y.cpu() != ref).sum().item()
    print(name, "OK" if ok else "MISMATCH", msg, f"{nd}/{ref.numel()} values differ bitwise", flush=True)
    if not ok:  # where: raw conv outputs first
        for l in range(3):
            print("   level", l, "max |y - ref| box", (y.cpu()[:, :4].float() - ref[:, :4].float()).abs().max().item(),
                  "cls", (y.cpu()[:, 4:].float() - ref[:, 4:].float()).abs().max().item())

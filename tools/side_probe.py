"""What do the delivery kernels on the third branch of the step graph cost?  One GPU plays both sides: the "writer"
pushes every batch from a staging buffer into a second LOCAL buffer with cerb_deliver_push and then acknowledges it to
itself with cerb_deliver_collect (world = 2, dst = 0, "rank 1" = this GPU), so the kernels, their fences and the graph
shape are exactly the multi-GPU ones and only NVLink is missing.

    python tools/side_probe.py [--steps 400]

Prints ms/step of the overlapped pipeline without a delivery and with the stand-alone kernels on a third branch of the
step graph, with and without their fences / copy and with 4 / 32 / 128 push CTAs.  Result (profiles/r02_multi_gpu.md):
+13 us per step whatever the kernels do -- the third branch itself costs the decode / NMS overlap -- which is why the
shipped delivery rides on the NMS launches instead (cerb_nms_deliver, piggyback form)."""
import argparse
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CERB_SIDE", "branch3")  # this probe is about the stand-alone kernels (read by pipeline.py at import)
import torch  # noqa: E402

from cerberusdet_b200 import _lib  # noqa: E402
from cerberusdet_b200.pipeline import PostHeadPipeline  # noqa: E402
from cerberusdet_b200.synth import STRIDES, synth_heads  # noqa: E402

NCS = [20, 19, 12]
KW = dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)


class LoopbackDelivery:
    """shard.PeerDelivery's interface with both ends on this GPU."""

    in_graph, direct, rank, dst = True, False, 1, 0

    def __init__(self, T, B, max_det, dev, mode):
        self.mode = mode  # "full" | "nofence": see push()
        n_d, n_c = T * B * max_det * 6, T * B
        self.n = (n_d + n_c + 63) // 64 * 64
        self.stage = [torch.zeros(self.n, device=dev) for _ in range(2)]
        self.remote = [torch.zeros(self.n, device=dev) for _ in range(2)]
        self.words = torch.zeros(64, dtype=torch.int32, device=dev)  # flags[slot][r] at 16*slot + r, ack[slot] at 32 + slot
        self.state = torch.zeros(8, dtype=torch.int32, device=dev)   # seq[2], done[2], collected[2]
        self.outs = [(s[:n_d].view(T, B, max_det, 6), s[n_d : n_d + n_c].view(torch.int32).view(T, B)) for s in self.stage]
        self.dev = dev

    def nms_deliver_args(self, slot, piggyback=False):
        return None

    def push(self, slot):
        lib = _lib.load()
        w, st = self.words.data_ptr(), self.state.data_ptr()
        s = torch.cuda.current_stream(self.dev).cuda_stream
        if self.mode != "collect_only":
            _lib.check(lib.cerb_deliver_push(self.stage[slot].data_ptr(), self.remote[slot].data_ptr(), self.n, w + 4 * (16 * slot + 1),
                                             w + 4 * (32 + slot), st + 4 * slot, st + 4 * (2 + slot), s))
        if self.mode == "memcpy":  # the same bytes through a copy node instead of a kernel
            self.remote[slot].copy_(self.stage[slot], non_blocking=True)
        if self.mode in ("full", "collect_only"):
            acks = _lib.ptr_array([None, w + 4 * (32 + slot)])
            _lib.check(lib.cerb_deliver_collect(w + 4 * 16 * slot, acks, st + 4 * (4 + slot), 2, 0, s))

    def collect(self, slot):
        pass


def run(steps, delivery_mode):
    dev = torch.device("cuda", 0)
    heads = [[x.to(dev) for x in lv] for lv in synth_heads(range(64), NCS, 640, torch.float16, "iid", cfg=3)]
    dv = LoopbackDelivery(len(NCS), 64, KW["max_det"], dev, delivery_mode) if delivery_mode not in ("none", "none_pad", "none_outs") else None
    pad = None
    outs = None
    if delivery_mode == "none_pad":    # the same allocations as LoopbackDelivery, unused: does buffer placement alone change the step time?
        pad = LoopbackDelivery(len(NCS), 64, KW["max_det"], dev, "full")
    if delivery_mode == "none_outs":   # ... or writing the detections into one packed buffer per slot?
        pad = LoopbackDelivery(len(NCS), 64, KW["max_det"], dev, "full")
        outs = pad.outs
    pipe = PostHeadPipeline(heads, STRIDES, KW, delivery=dv, outs=outs)

    def loop(n):
        pipe.k, pipe.pending = 0, None
        for _ in range(n):
            pipe.step()
        pipe.flush()

    loop(20)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loop(steps)
    e1.record()
    torch.cuda.synchronize()
    ok = True
    if dv is not None:
        last = (steps - 1) & 1
        ok = bool(torch.equal(dv.remote[last], dv.stage[last]))
    return e0.elapsed_time(e1) / steps, ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--one", default=None)
    args = ap.parse_args()
    if args.one:
        ms, ok = run(args.steps, args.one)
        print(json.dumps({"mode": args.one, "side": os.environ.get("CERB_SIDE", "branch"), "push_mode": os.environ.get("CERB_DEBUG_PUSH_MODE"),
                          "push_ctas": os.environ.get("CERB_DEBUG_PUSH_CTAS"), "schedule": os.environ.get("CERB_SCHEDULE", "pdl"), "ms_per_step": round(ms, 5), "delivered": ok}))
        return
    # (a collect without a matching flag-setting push would sit out its 2 s bound every step: never combine those)
    if os.environ.get("SIDE_PROBE_SET", "placement") == "dummy":  # is it buffer placement, the packed outputs, or the side kernel?  (bimodal step times)
        runs = (("none", {}), ("none_pad", {}), ("none_outs", {}), ("full", {"CERB_SIDE": "last3"}), ("full", {"CERB_SIDE": "last3", "CERB_DEBUG_PUSH_MODE": "2"})) * 2
    elif os.environ.get("SIDE_PROBE_SET", "placement") == "placement":  # where in the step graph does the side kernel hurt least?
        runs = (("none", {}),) + tuple(("full", {"CERB_SIDE": v}) for v in ("branch3", "last3", "pre_nms", "post_nms", "post_decode")) + (("none", {}),)
    else:  # what inside the side kernels costs? (a collect without a matching flag-setting push would sit out its 2 s bound every step)
        runs = (("none", {}), ("full", {}), ("push_only", {"CERB_DEBUG_PUSH_MODE": "1"}), ("full", {"CERB_DEBUG_PUSH_MODE": "2"}),
                ("full", {"CERB_DEBUG_PUSH_CTAS": "4"}), ("full", {"CERB_DEBUG_PUSH_CTAS": "128"}), ("none", {}), ("full", {}))
    for mode, extra in runs:
        env = dict(os.environ, **extra)
        try:
            out = subprocess.run([sys.executable, __file__, "--steps", str(args.steps), "--one", mode], env=env, capture_output=True,
                                 text=True, timeout=90)
            print(out.stdout.strip() or out.stderr[-400:], flush=True)
        except subprocess.TimeoutExpired:
            print(json.dumps({"mode": mode, "extra": extra, "error": "timeout"}), flush=True)


if __name__ == "__main__":
    main()

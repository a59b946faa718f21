#!/usr/bin/env python
"""bench.py -- post-head path throughput (decode + per-task NMS) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one pass of the hot path over one batch of synthetic raw head tensors:
BASELINE.json config 3 (3 task heads 20/19/12 classes, 640x640 -> 8400 anchors,
B=64 images per GPU, fp16, val settings conf 0.001 / iou 0.6 / multi_label /
max_nms 30000 / max_det 300).  N>1: every rank owns B images (weak scaling; N=8 is
config 5, B=512) and the padded detections are gathered to rank 0 inside the step.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` goes
through the host-buffer public API (H2D of the raw heads and D2H of the detections
inside the timed region), `roofline` is the decode kernel against the measured HBM
peak, `cpu_baseline` is the oracle port (the reference's torch/torchvision algorithm)
timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

NCS = [20, 19, 12]  # voc / objects365_animals / objects365_tableware (reference data/*.yaml)
IMGSZ = 640
B_PER_GPU = 64
NMS_KW = dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)  # reference val.py:139,318
METRIC = "post-proc images/s (3 tasks, 640^2, B=64 per GPU): Detect decode + per-task NMS"
CPU_SAMPLE_IMAGES = 3  # cpu_baseline: ~1.2 s per (image, task) segment on 16 cores -> ~10 s of CPU work
EVENT_EVERY = 8  # graph mode: every 8th timed step brackets the decode kernel with events (roofline sample)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _decode_bytes_per_image(ncs, anchors, elt):
    return sum((64 + nc) * elt * anchors + (4 + nc) * elt * anchors for nc in ncs)


class ClockSampler:
    """SM clock + throttle reasons of one GPU while the timed region runs (NVML, ~2 ms period;
    falls back to polling nvidia-smi)."""

    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.max_mhz = index, [], set(), None
        self.nvml, self.h = None, None
        self.stop = threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)

    def _init_nvml(self):
        """Open the NVML handle on the caller's thread, BEFORE the timed region: nvmlInit alone can take longer than
        the ~50 ms the default run is timed for, which would leave the sampler without a single sample."""
        import pynvml

        pynvml.nvmlInit()
        # NVML enumerates physical GPUs; honour CUDA_VISIBLE_DEVICES when it is a plain index list
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = self.index
        if vis:
            try:
                phys = int(vis.split(",")[self.index])
            except (ValueError, IndexError):
                phys = self.index
        self.nvml, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))

    def _sample_nvml(self):
        self.sm.append(float(self.nvml.nvmlDeviceGetClockInfo(self.h, self.nvml.NVML_CLOCK_SM)))
        mask = int(self.nvml.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        for name, bit in self.BAD.items():
            if mask & bit:
                self.reasons.add(name)

    def _run_nvml(self):
        while not self.stop.is_set():
            self._sample_nvml()
            self.stop.wait(0.002)
        self._sample_nvml()  # one more under load: the caller stops the sampler before it synchronises

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.sm.append(float(parts[0]))
                    self.max_mhz = float(parts[1])
                    self.reasons.update(n for i, n in enumerate(names) if parts[2 + i].lower().startswith("active"))
            except Exception:
                pass
            self.stop.wait(0.05)

    def _run(self):
        try:
            if self.h is None:
                raise RuntimeError("no NVML")
            self._run_nvml()
        except Exception:
            self._run_smi()

    def __enter__(self):
        try:
            self._init_nvml()
        except Exception:
            self.h = None
        self.th.start()
        return self

    def __exit__(self, *exc):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"], "samples": 0}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


def cpu_port_segment(heads_cpu, image, task):
    """The oracle port (reference algorithm on torch CPU) on ONE (image, task) segment:
    decode of that head, then non_max_suppression with one image per call."""
    from cerberusdet_b200.synth import STRIDES
    from oracle import ref_port as rp

    t0 = time.perf_counter()
    lv = [x[image : image + 1] for x in heads_cpu[task]]
    y = rp.decode_port(lv, NCS[task], STRIDES)
    rp.nms_port(y, greedy="torchvision", **NMS_KW)
    return time.perf_counter() - t0


REFERENCE_BUDGET_S = 240.0


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port: same torch ops +
    torchvision.ops.nms) on the host cores.  Each step is a bounded sample of the workload:
    one (image, task) segment = 1/3 image, the task rotating with the step."""
    if rank != 0:
        return
    from cerberusdet_b200.synth import synth_heads

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    heads = synth_heads(range(CPU_SAMPLE_IMAGES), NCS, IMGSZ, torch.float16, "iid", cfg=3)
    t_begin = time.perf_counter()
    for k in range(args.warmup):
        cpu_port_segment(heads, 0, k % 3)
    times = []
    for k in range(args.steps):
        times.append(cpu_port_segment(heads, 0, k % 3))
        if time.perf_counter() - t_begin > REFERENCE_BUDGET_S and len(times) >= 3:
            break  # keep the arm inside a few minutes whatever K the caller asked for
    tot, done = sum(times), len(times)
    val = (done / 3.0) / tot
    sample = (f"one (image, task) segment (= 1/3 image) of the B=64 batch per step, task rotating; "
              f"{done} of {args.steps} steps run")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * tot / done,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": _config(args.gpus),
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _config(n):
    return {"workload": "BASELINE config 3 per GPU: 3 task heads (nc 20/19/12), 640x640 (8400 anchors), "
                        f"B={B_PER_GPU} images per GPU, fp16 raw head tensors, conf 0.001 / iou 0.6 / multi_label / "
                        "max_nms 30000 / max_det 300",
            "global_batch": B_PER_GPU * n, "tasks": 3, "anchors": 8400, "regime": "iid (SURVEY 8d R-dense)",
            "l2": "inputs (261 MB raw heads per step) exceed the 126 MB L2; no explicit flush",
            "parallelism": f"image-sharded x{n}, gather of padded detections to rank 0" if n > 1 else "single GPU"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="launch the two kernels eagerly instead of replaying CUDA graphs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist

    from cerberusdet_b200 import ops
    from cerberusdet_b200.api import postprocess_host
    from cerberusdet_b200.shard import DetectionGatherer
    from cerberusdet_b200.synth import STRIDES, synth_heads

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    images = range(rank * B_PER_GPU, (rank + 1) * B_PER_GPU)
    heads_host = synth_heads(images, NCS, IMGSZ, torch.float16, "iid", cfg=3, pin=True)
    heads_dev = [[x.to(dev, non_blocking=True) for x in lv] for lv in heads_host]
    torch.cuda.synchronize()

    # N > 1: the NMS kernel writes into a packed per-rank buffer that ONE asynchronous gather moves to
    # rank 0; two buffers alternate so the gather of step k overlaps the kernels of step k+1.
    gatherers = [DetectionGatherer(len(NCS), B_PER_GPU, NMS_KW["max_det"], dev, dst=0) for _ in range(2)] if world > 1 else None
    step_no = [0]

    def step(record=None):
        if record is not None:
            record[0].record()
        ys = ops.decode_heads(heads_dev, STRIDES)
        if record is not None:
            record[1].record()
        if world > 1:
            g = gatherers[step_no[0] & 1]
            step_no[0] += 1
            g.wait()  # the gather issued from this buffer two steps ago
            dets, counts = ops.nms_batched(ys, out=g.out, **NMS_KW)
            if record is not None:
                record[2].record()
            g.launch()
        else:
            dets, counts = ops.nms_batched(ys, **NMS_KW)
            if record is not None:
                record[2].record()
        return dets, counts

    def drain():
        if world > 1:
            for g in gatherers:
                g.wait()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup // 2, 3)):
        step()
    drain()
    barrier()

    # Launch-bound inner loop -> CUDA graphs.  A plain step replays ONE graph (decode kernel -> NMS kernel); an
    # instrumented step replays graph A = decode kernel and graph B = NMS kernel separately so that the decode kernel
    # can be bracketed by events inside the timed region.  (The gather to rank 0 for N > 1 stays outside the graphs.)
    mode = "eager"
    if not args.no_graphs:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                g_dec = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_dec, stream=side):
                    ys_static = ops.decode_heads(heads_dev, STRIDES)
                # one NMS graph per output buffer (N > 1 alternates two gather buffers); the NCCL gather itself
                # stays outside the graphs and is issued asynchronously after the replay
                g_nms, outs_static = [], []
                for i in range(2 if world > 1 else 1):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        o = ops.nms_batched(ys_static, out=gatherers[i].out, **NMS_KW) if world > 1 else ops.nms_batched(ys_static, **NMS_KW)
                    g_nms.append(g)
                    outs_static.append(o)
                # the same two kernels as ONE graph (decode -> NMS), for the steps whose decode kernel is not bracketed
                # by events: two graph launches + an event record between them leave ~5 us of idle GPU per kernel
                g_full, outs_full = [], []
                for i in range(2 if world > 1 else 1):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        ys_f = ops.decode_heads(heads_dev, STRIDES)
                        o = ops.nms_batched(ys_f, out=gatherers[i].out, **NMS_KW) if world > 1 else ops.nms_batched(ys_f, **NMS_KW)
                    g_full.append(g)
                    outs_full.append(o)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()

            def step(record=None):  # noqa: F811
                i = step_no[0] & 1 if world > 1 else 0
                step_no[0] += 1
                if record is None:  # plain step: one graph launch
                    if world > 1:
                        gatherers[i].wait()  # the gather issued from this buffer two steps ago
                    g_full[i].replay()
                    if world > 1:
                        gatherers[i].launch()
                    return outs_full[i]
                record[0].record()
                g_dec.replay()
                record[1].record()
                if world > 1:
                    gatherers[i].wait()
                g_nms[i].replay()
                record[2].record()
                if world > 1:
                    gatherers[i].launch()
                return outs_static[i]

            mode = "cuda_graphs"
        except Exception as exc:  # pragma: no cover - depends on the driver/NCCL build
            print(f"[bench] CUDA graph capture failed ({type(exc).__name__}: {exc}); running eagerly", file=sys.stderr)
            torch.cuda.synchronize()
    for _ in range(args.warmup):
        step()
    drain()
    barrier()
    # the decode kernel is bracketed by events on every EVENT_EVERY-th step of the timed region (the instrumented
    # step replays the decode graph and the NMS graph separately); eager mode brackets every step
    every = EVENT_EVERY if mode == "cuda_graphs" else 1
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] if k % every == 0 else None for k in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        t_start.record()
        for k in range(args.steps):
            step(evs[k])
        drain()  # every batch's detections have reached rank 0
        t_end.record()
        barrier()
    elapsed_ms = t_start.elapsed_time(t_end)
    dec_ms = [e[0].elapsed_time(e[1]) for e in evs if e is not None]
    nms_ms = [e[1].elapsed_time(e[2]) for e in evs if e is not None]
    if world > 1:
        tt = torch.tensor([elapsed_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tt.item())

    # ---- end to end through the host-buffer API (pinned host in, pinned host out)
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        postprocess_host(heads_host, STRIDES, device=dev, **NMS_KW)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out_dets, out_counts = postprocess_host(heads_host, STRIDES, device=dev, **NMS_KW)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    h2d = sum(x.numel() * x.element_size() for lv in heads_host for x in lv)
    d2h = out_dets.numel() * 4 + out_counts.numel() * 4

    if rank == 0:
        peak, peak_src = _peaks()
        elt = 2
        bytes_per_launch = B_PER_GPU * _decode_bytes_per_image(NCS, 8400, elt)
        dec_avg_ms = sum(dec_ms) / len(dec_ms)
        achieved = bytes_per_launch / (dec_avg_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "decode_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": B_PER_GPU * world * args.steps / (elapsed_ms * 1e-3), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": _config(world),
            "roofline": {"kernel": "decode_pipe_kernel<__half,8,2>", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "frac_of_nominal_8000": achieved / 8000.0, "algorithmic_bytes_per_launch": bytes_per_launch,
                         "decode_ms_avg": dec_avg_ms, "decode_ms_median": statistics.median(dec_ms),
                         "nms_ms_avg": sum(nms_ms) / len(nms_ms), "nms_ms_median": statistics.median(nms_ms), "timed_launches": len(dec_ms)},
            "e2e": {"value": B_PER_GPU * world * e2e_steps / e2e_s, "unit": "images/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": 2 * args.steps,  # decode_pipe_kernel + nms_kernel per step, nothing else
            "launch_mode": mode,
            "clocks": clk.summary(),
        }
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            cpu_heads = [[x[:CPU_SAMPLE_IMAGES].clone() for x in lv] for lv in heads_host]
            cpu_port_segment(cpu_heads, 0, 0)  # warm-up
            secs = sum(cpu_port_segment(cpu_heads, i, t) for i in range(CPU_SAMPLE_IMAGES) for t in range(3))
            line["cpu_baseline"] = {"value": CPU_SAMPLE_IMAGES / secs, "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": f"images 0..{CPU_SAMPLE_IMAGES - 1} of the same batch, all 3 task heads, once "
                                              f"({secs:.1f} s; oracle port: torch CPU ops + torchvision.ops.nms)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Streaming engine of the post-head path on one GPU: static buffers, CUDA graphs, and a two-stage software pipeline
in which the decode of batch k overlaps the NMS of batch k-1.

Why overlap.  The decode kernel is memory-bound (all SMs streaming 330 MB per config-3 batch); the NMS kernel is a
latency-bound chain of short phases that keeps the SMs ~30 % busy and touches little memory.  Run back to back they
take 64 + 52 us; run concurrently -- NMS of the previous batch on a second stream while the next batch is decoded --
a step takes ~98 us (profiles/r02_nms.md, "The overlapped step").  Both kernels are the same launches as in the serial path; only their
placement in time changes, so results are bit-identical to ``ops.decode_heads`` + ``ops.nms_batched``.

Step k (one CUDA-graph replay):      stream A:  decode(inputs)      -> Y[k & 1]
                                     stream B:  NMS(Y[(k-1) & 1])   -> OUT[(k-1) & 1]
``flush()`` runs the NMS of the last decoded batch.  Inputs are static device tensors (a serving loop copies -- or
lets the conv towers write -- each batch into them before ``step()``); outputs are static ``(dets, counts)`` buffers,
which may be caller-provided (e.g. ``shard.DetectionGatherer`` buffers, or rank 0's peer-mapped memory).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import ops


class PostHeadPipeline:
    def __init__(self, heads: Sequence[Sequence[torch.Tensor]], strides: Sequence[float], nms_kw: dict,
                 outs: Optional[Sequence[Tuple[torch.Tensor, torch.Tensor]]] = None, timed_parities: Sequence[int] = (),
                 overlap: bool = True, delivery=None):
        """``heads[t][l]``: static raw head tensors ``[B, 64+nc_t, H_l, W_l]`` on one CUDA device.  ``outs``: two
        ``(dets[T,B,max_det,6] float32, counts[T,B] int32)`` buffer pairs (allocated here when omitted).
        ``timed_parities`` additionally captures one instrumented step graph per entry (the NMS kernel, then the decode
        kernel alone, each between timing events recorded by the graph itself) for ``step(timed=i)``; entry i is the
        parity (step index & 1) of the step slot i will be used at.  ``overlap=False`` captures the serial
        order (decode -> NMS of the same batch, one stream) behind the same interface.
        ``delivery`` (N > 1): a ``shard.PeerDelivery`` whose hand-shake runs inside the kernels (``in_graph``): the NMS
        launches carry its protocol words and rank dst's ``collect`` of the batch delivered one step earlier is captured
        in the same step graph, so a step stays ONE graph replay on every rank."""
        first = heads[0][0]
        if not first.is_cuda:
            raise TypeError("PostHeadPipeline needs CUDA tensors (cerberusdet_b200 has no CPU path)")
        self.device = first.device
        self.heads, self.strides, self.kw = [list(lv) for lv in heads], [float(s) for s in strides], dict(nms_kw)
        self.overlap = bool(overlap)
        self.delivery = delivery if (delivery is not None and getattr(delivery, "in_graph", False)) else None
        if self.delivery is not None:
            if outs is None:
                outs = self.delivery.outs
            if not self.overlap:
                raise ValueError("in-graph delivery is wired for the overlapped schedule")
        T, B = len(heads), int(first.shape[0])
        max_det = int(self.kw.get("max_det", 300))
        if outs is None:
            outs = [(torch.empty((T, B, max_det, 6), dtype=torch.float32, device=self.device),
                     torch.empty((T, B), dtype=torch.int32, device=self.device)) for _ in range(2)]
        if len(outs) != 2:
            raise ValueError("outs must hold exactly two (dets, counts) buffer pairs")
        self.outs = list(outs)
        self.k = 0            # steps issued in this run (a run = k = 0 .. n-1, then flush(); set k = 0, pending = None to start one)
        self.pending = None   # parity of the decoded batch whose NMS has not been issued yet
        with torch.cuda.device(self.device):
            self.ybuf = [ops.decode_buffers(self.heads) for _ in range(2)]
            self.sa, self.sb = torch.cuda.Stream(), torch.cuda.Stream()
            # eager warm-up of both kernels (module load, function attributes) before anything is captured
            ys = ops.decode_heads(self.heads, self.strides, out=self.ybuf[0])
            ops.nms_batched(ys, out=self.outs[0], **self.kw)
            torch.cuda.synchronize(self.device)
            self._g_first = [self._capture(p, nms_of=None) for p in (0, 1)]              # decode only (first step)
            self._g_step = [self._capture(p, nms_of=p if not self.overlap else 1 - p, collect=True) for p in (0, 1)]
            self._g_flush = [self._capture(None, nms_of=p, collect=True) for p in (0, 1)]              # NMS only (drain)
            if self.delivery is not None:
                # second step of a run: nothing has been delivered into the other slot yet, so nothing to collect;
                # a run of one step: only the flushed batch is collected; after an instrumented step: collect alone
                self._g_step1 = self._capture(1, nms_of=0, collect=False)
                self._g_flush1 = self._capture(None, nms_of=0, collect="own")
                self._g_collect = [self._capture_collect(p) for p in (0, 1)]
            self.timed: List[Tuple[torch.cuda.CUDAGraph, list]] = []
            self.timed_parity = [int(p) & 1 for p in timed_parities]
            for p in self.timed_parity:
                self.timed.append(self._capture_timed(p))
            torch.cuda.synchronize(self.device)

    # ------------------------------------------------------------------ graph construction
    def _decode(self, p):
        return ops.decode_heads(self.heads, self.strides, out=self.ybuf[p])

    def _nms(self, p):
        T = len(self.heads)
        ys = self.ybuf[p][:T]
        for y, sm in zip(ys, self.ybuf[p][T:]):  # the summary of a static buffer describes whatever decode wrote last
            if sm.shape[-1]:
                ops._remember_summary(y, sm)
        deliver = self.delivery.nms_deliver_args(p) if self.delivery is not None else None
        return ops.nms_batched(ys, out=self.outs[p], deliver=deliver, **self.kw)

    def _capture_collect(self, slot: int):
        if self.delivery.rank != self.delivery.dst:
            return None  # collect() is dst's side only
        g = torch.cuda.CUDAGraph()
        self.sa.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.sa):
            with torch.cuda.graph(g, stream=self.sa):
                self.delivery.collect(slot)
        torch.cuda.current_stream(self.device).wait_stream(self.sa)
        return g

    def _capture(self, dec: Optional[int], nms_of: Optional[int], collect=False):
        """``collect`` (in-graph delivery, rank dst): True = after the NMS also collect the batch delivered one step
        earlier (the other slot), and for a flush both; "own" = only the batch this NMS delivers."""
        dv = self.delivery
        g = torch.cuda.CUDAGraph()
        self.sa.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.sa):
            with torch.cuda.graph(g, stream=self.sa):
                if self.overlap or dec is None or nms_of is None:
                    if nms_of is not None and dec is not None:
                        self.sb.wait_stream(self.sa)
                        with torch.cuda.stream(self.sb):
                            self._nms(nms_of)
                            if dv is not None and collect:
                                dv.collect(1 - nms_of)  # the batch the other ranks delivered during the previous step
                        self._decode(dec)
                        self.sa.wait_stream(self.sb)
                    elif dec is not None:
                        self._decode(dec)
                    else:
                        self._nms(nms_of)
                        if dv is not None and collect is True:
                            dv.collect(1 - nms_of)
                        if dv is not None and collect:
                            dv.collect(nms_of)
                else:  # serial: decode -> NMS of the SAME batch (the NMS kernel starts under programmatic dependent launch)
                    self._decode(dec)
                    self._nms(nms_of)
        torch.cuda.current_stream(self.device).wait_stream(self.sa)
        return g

    def _capture_timed(self, p: int):
        """Instrumented step: the same two launches as a normal step (decode of this batch, NMS of the previous one) in
        SERIAL order with timing events recorded by graph nodes:  E0 ; decode ; {E1 on a side branch} ; NMS ; E2.
        E1 hangs off the decode kernel on a second stream, so the NMS kernel keeps its programmatic (early-launch) edge
        to the decode kernel; E0..E1 is the decode kernel alone on the GPU, E1..E2 the NMS kernel."""
        ev = [torch.cuda.Event(enable_timing=True, external=True) for _ in range(3)]
        g = torch.cuda.CUDAGraph()
        self.sa.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.sa):
            with torch.cuda.graph(g, stream=self.sa):
                ev[0].record(self.sa)
                self._decode(p)
                self.sb.wait_stream(self.sa)
                ev[1].record(self.sb)
                self._nms(1 - p)
                ev[2].record(self.sa)
                self.sa.wait_stream(self.sb)
        torch.cuda.current_stream(self.device).wait_stream(self.sa)
        return g, ev

    # ------------------------------------------------------------------ running
    def step(self, timed: Optional[int] = None) -> Optional[int]:
        """Issue one pipeline step on the current stream.  Returns the index of the ``outs`` buffer that this step's
        NMS fills (the detections of the PREVIOUS batch; of this batch when ``overlap=False``), or None on the first
        step of an overlapped pipeline."""
        k = self.k
        p = k & 1
        self.k += 1
        if not self.overlap:
            self._g_step[p].replay()
            return p
        if self.pending is None:
            self._g_first[p].replay()
            self.pending = p
            return None
        if timed is not None:
            g, _ = self.timed[timed]
            if self.timed_parity[timed] != p:
                raise ValueError("timed slot parity does not match the step parity")
            g.replay()
            if self.delivery is not None and k >= 2 and self._g_collect[p] is not None:
                self._g_collect[p].replay()  # (the instrumented graph holds the two kernels and their events only)
        elif self.delivery is not None and k < 2:
            self._g_step1.replay()           # second step of a run: no batch in the other slot yet
        else:
            self._g_step[p].replay()
        done, self.pending = self.pending, p
        return done

    def flush(self) -> Optional[int]:
        """NMS of the last decoded batch (end of a stream of batches).  Returns its ``outs`` index."""
        if not self.overlap or self.pending is None:
            return None
        p, self.pending = self.pending, None
        if self.delivery is not None and self.k < 2:
            self._g_flush1.replay()  # a run of one step: only the flushed batch is there to collect
        else:
            self._g_flush[p].replay()
        return p

    def timed_ms(self, i: int) -> Tuple[float, float]:
        """(NMS ms, decode ms) of the last replay of instrumented slot ``i`` (synchronise first)."""
        _, ev = self.timed[i]
        return ev[1].elapsed_time(ev[2]), ev[0].elapsed_time(ev[1])

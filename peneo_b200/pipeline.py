"""Streaming heads + decode: the decoder-side loop of deploy/inference.py (``inference()`` ->
``postprocess()``, deploy/inference.py:375-386, 388-462) with the GPU and the host overlapped.

``submit()`` enqueues, without blocking, the H2D copy of a batch of hidden states (copy stream), the
per-token projections, pair heads, spot extraction and link resolution (compute stream) and the D2H copy
of the compact records; ``result()`` waits for that batch only and builds the reference's Python
objects while the GPU already works on the next batch.  Documents are independent, so this is also the
unit that is sharded across GPUs (one pipeline per process / GPU, no collective).
"""
from __future__ import annotations

import time
from collections import deque
from typing import List, Optional, Sequence

import torch

from . import _lib, decode, ops


class _GraphSlot:
    """One in-flight batch with every device buffer static: the three per-token GEMMs, the pair kernel, spot extraction,
    link resolution and the D2H copies of the compact records are captured ONCE into a CUDA graph and replayed per
    batch — one launch instead of seven plus three memsets and two copies, no launch gaps, ~0.25 ms less host work per
    step.  The H2D (or D2D) copy of the hidden states into ``x`` stays outside the graph."""

    def __init__(self, pipe: "HeadsDecodePipeline", b: int, n: int, hin: int, dtype: torch.dtype):
        dev, dec = pipe.device, pipe.decoder
        self.key = (b, n, hin, dtype)
        self.x = torch.empty(b, n, hin, dtype=dtype, device=dev)
        lib = _lib.load()
        p = ops.shaking_len(n)
        cap = min(p, max(8 * n, 1024))
        doc_ints = lib.peneo_decode_resolve_doc_ints(n, cap)
        host = (torch.empty(b * decode.NUM_HEADS, dtype=torch.int32, pin_memory=True),
                torch.empty(b, doc_ints, dtype=torch.int32, pin_memory=True))
        ext = lambda: torch.cuda.Event(enable_timing=True, external=True)  # noqa: E731  (timed inside the graph)
        self.k2_ev = (ext(), ext())
        self.done = torch.cuda.Event()
        self.kernels_done = torch.cuda.Event()
        self.graph = torch.cuda.CUDAGraph()
        k0 = ops.COUNTERS["kernels"]
        with torch.cuda.stream(pipe.compute), torch.no_grad():
            pack = dec._weight_pack(dev)
            self._run(pipe, pack, b, n, cap, host, None)  # eager warm-up: function attributes, driver entry points, allocator
        pipe.compute.synchronize()
        with torch.no_grad(), torch.cuda.graph(self.graph, stream=pipe.compute):
            self.pending = self._run(pipe, pack, b, n, cap, host, (self.k2_ev,))
        self.kernels = (ops.COUNTERS["kernels"] - k0) // 2
        self.d2h_bytes = self.pending.d2h_bytes

    def _run(self, pipe, pack, b, n, cap, host, ev):
        ab = ops.token_projections(pack, self.x)
        if pipe.fused_spots:
            if ev:
                ev[0][0].record(pipe.compute)
            pending = decode.PendingDecode(None, n, cap, False, pipe.score_thresh, False, None, None, (pack, ab), host, False,
                                           defer_d2h=True)
            if ev:
                ev[0][1].record(pipe.compute)
            return pending
        if ev:
            ev[0][0].record(pipe.compute)
        logits = ops.pair_heads(pack, ab, b, n)
        if ev:
            ev[0][1].record(pipe.compute)
        ins, _ = decode._as_batched_inputs(logits)
        pending = decode.PendingDecode(ins, n, cap, False, pipe.score_thresh, False, None, None, None, host, False,
                                       defer_d2h=True)
        return pending


class HeadsDecodePipeline:
    def __init__(self, decoder, device=None, score_thresh: float = 0.0, fused_spots: Optional[bool] = None,
                 use_graphs: Optional[bool] = None):
        self.decoder = decoder
        self.device = torch.device(device) if device is not None else next(decoder.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("HeadsDecodePipeline needs the decoder on a CUDA device (no CPU path)")
        self.compute = torch.cuda.Stream(self.device)
        self.copy = torch.cuda.Stream(self.device)
        self.d2h = torch.cuda.Stream(self.device)
        self.score_thresh = score_thresh
        # `fused_spots=True`: spot extraction inside the pair kernel (no [B, P, C] logits written or read, 234 MB per
        # 32 x seq-512 batch less HBM traffic and memory).  Measured on B200 it is 1.3 % SLOWER end to end than
        # K2 -> logits -> K3 (6 260 vs 6 340 docs/s in 40-step runs): K2 is paced by its CUDA-core warps, the classify /
        # ballot work costs them more than the 0.09 ms the stand-alone K3 takes.  So the default is the logits route;
        # the fused route is for memory-bound deployments (N = 2048: 117 MB of logits per document).
        fusable = decoder.precision == "bf16" and decoder.dims.bf16_capable()
        if fused_spots and not fusable:
            raise ValueError("fused_spots needs the fused tcgen05 configuration (precision 'bf16', shrink, 768 -> 384, 2 layers)")
        if fused_spots is None:  # PENEO_FUSED_SPOTS=0/1 overrides the default (A/B studies)
            import os

            env = os.environ.get("PENEO_FUSED_SPOTS")
            fused_spots = False if env is None else (fusable and env != "0")
        self.fused_spots = bool(fused_spots)
        # CUDA graphs (one per in-flight slot and batch shape) unless PENEO_GRAPHS=0 / use_graphs=False
        if use_graphs is None:
            import os

            use_graphs = os.environ.get("PENEO_GRAPHS", "1") != "0"
        self.use_graphs = bool(use_graphs)
        self._free_slots = {}   # (b, n, hin, dtype) -> idle _GraphSlot objects
        self._shape_seen = {}   # (b, n, hin, dtype) -> submissions so far (shapes without a graph yet)
        self.max_graph_shapes = 8
        self.kernel_ms = None   # graph mode: set to {"k2": []} to collect K2's per-launch time (external CUDA events in the graph)
        self._queue = deque()
        self.h2d_bytes = 0
        self.wait_s = 0.0      # result(): time blocked on the GPU
        self.assemble_s = 0.0  # result(): time building the Python result objects
        self.d2h_bytes = 0
        self.k2_events = None  # set to a list to collect (start, stop) CUDA events around K2
        self.k3_events = None  # same for the spot-extraction kernel K3

    def submit(self, hidden: torch.Tensor, texts: Sequence[List[str]], bboxes=None):
        """hidden: [B, N, Hin] on the host (pinned for a truly asynchronous copy) or on the device."""
        b, n, _ = hidden.shape
        if self.use_graphs and self._graph_worthwhile(hidden):
            return self._submit_graph(hidden, texts, bboxes)
        if hidden.is_cuda:
            x = hidden
            ready = torch.cuda.Event()  # whatever produced `hidden` on the caller's stream must be done first
            ready.record(torch.cuda.current_stream(self.device))
        else:
            with torch.cuda.stream(self.copy):
                x = hidden.to(self.device, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(self.copy)
            self.h2d_bytes += hidden.numel() * hidden.element_size()
        with torch.cuda.stream(self.compute), torch.no_grad():
            if ready is not None:
                self.compute.wait_event(ready)
            pack = self.decoder._weight_pack(self.device)
            ab = ops.token_projections(pack, x)
            if self.fused_spots:
                # K2 with the spots-only epilogue + gather (timed together as "K2": the gather is ~10 us)
                pending = decode.heads_decode_async(pack, ab, b, n, score_thresh=self.score_thresh, d2h_stream=self.d2h,
                                                    k2_events=self.k2_events)
                self.decoder._pack_read_done(self.device)
            else:
                if self.k2_events is not None:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(self.compute)
                logits = ops.pair_heads(pack, ab, b, n)
                if self.k2_events is not None:
                    e1.record(self.compute)
                    self.k2_events.append((e0, e1))
                self.decoder._pack_read_done(self.device)  # a later re-pack on another stream waits for these kernels
                pending = decode.device_decode_async(logits, n, score_thresh=self.score_thresh, d2h_stream=self.d2h,
                                                     k3_events=self.k3_events)
            x.record_stream(self.compute)
        self.d2h_bytes += pending.d2h_bytes
        self._queue.append((pending, list(texts), bboxes))
        return len(self._queue)

    def _graph_worthwhile(self, hidden: torch.Tensor) -> bool:
        """A CUDA graph is captured per batch SHAPE (one eager warm-up + one capture, static buffers kept): worth it for a
        shape that keeps coming back (a serving loop with fixed batches), not for a stream of mixed lengths where most
        shapes occur once — measured on the 2 000-document mixed-length sweep: 228 docs/s with a graph per new shape
        against the eager path's several thousand.  A shape goes to graphs from its third occurrence on, and at most
        ``max_graph_shapes`` shapes hold graphs."""
        b, n, hin = hidden.shape
        dtype = hidden.dtype if hidden.dtype in (torch.float32, torch.bfloat16, torch.float16) else torch.float32
        key = (b, n, hin, dtype)
        if key in self._free_slots:
            return True
        seen = self._shape_seen.get(key, 0) + 1
        self._shape_seen[key] = seen
        if len(self._shape_seen) > 4096:  # (bounded bookkeeping for endless streams of new shapes)
            self._shape_seen.clear()
        return seen >= 3 and len(self._free_slots) < self.max_graph_shapes

    def _submit_graph(self, hidden: torch.Tensor, texts, bboxes):
        b, n, hin = hidden.shape
        dtype = hidden.dtype if hidden.dtype in (torch.float32, torch.bfloat16, torch.float16) else torch.float32
        key = (b, n, hin, dtype)
        free = self._free_slots.setdefault(key, [])
        slot = free.pop() if free else _GraphSlot(self, b, n, hin, dtype)
        if hidden.is_cuda:
            ready = torch.cuda.Event()  # whatever produced `hidden` on the caller's stream must be done first
            ready.record(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.compute):
                self.compute.wait_event(ready)
                slot.x.copy_(hidden, non_blocking=True)
                hidden.record_stream(self.compute)
        else:
            with torch.cuda.stream(self.copy):
                slot.x.copy_(hidden, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(self.copy)
            self.compute.wait_event(ready)
            self.h2d_bytes += hidden.numel() * hidden.element_size()
        with torch.cuda.stream(self.compute), torch.no_grad():
            self.decoder._weight_pack(self.device)  # re-packs (eagerly, same buffer) if a parameter changed
            slot.graph.replay()
            slot.kernels_done.record(self.compute)
            self.decoder._pack_read_done(self.device)  # a later re-pack on another stream waits for this replay
        # the record blocks go home on the D2H stream (4.6 MB per 32 x seq-512 batch: ~0.1 ms of PCIe time that would
        # otherwise sit between this batch's kernels and the next batch's in the compute stream)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(slot.kernels_done)
            slot.pending.d2h_copy()
            slot.done.record(self.d2h)
        ops.COUNTERS["kernels"] += slot.kernels
        slot.pending.d2h_bytes = slot.d2h_bytes
        self.d2h_bytes += slot.d2h_bytes
        self._queue.append((slot, list(texts), bboxes))
        return len(self._queue)

    def result(self, assemble: bool = True):
        """Python results (one reference 7-tuple per document) of the oldest submitted batch.
        ``assemble=False`` returns the raw :class:`decode.DeviceDecode` (records on the host; in graph mode they are
        views of the slot's pinned buffers, valid until that slot is submitted again)."""
        pending, texts, bboxes = self._queue.popleft()
        t0 = time.perf_counter()
        slot = pending if isinstance(pending, _GraphSlot) else None
        with torch.cuda.stream(self.compute):
            if slot is not None:
                slot.done.synchronize()
                if self.kernel_ms is not None:
                    self.kernel_ms["k2"].append(slot.k2_ev[0].elapsed_time(slot.k2_ev[1]))
                dd = slot.pending.finish()  # (re-runs documents whose spot lists overflowed, eagerly)
                self.d2h_bytes += slot.pending.d2h_bytes - slot.d2h_bytes
            else:
                dd = pending.finish()
        t1 = time.perf_counter()
        self.wait_s += t1 - t0  # host blocked on the GPU (0 when the host is the slower side)
        if not assemble:
            if slot is not None:
                self._free_slots[slot.key].append(slot)
            return dd
        out = decode.assemble_many(dd, range(dd.batch), texts, bboxes)
        if slot is not None:
            self._free_slots[slot.key].append(slot)
        self.assemble_s += time.perf_counter() - t1
        return out

    def __len__(self):
        return len(self._queue)

    @staticmethod
    def freeze_host_gc() -> None:
        """Serving-loop hygiene for the host side: ``result()`` creates tens of thousands of small containers per
        batch, which periodically triggers a full (generation-2) collection of the CPython cycle collector; with
        everything torch has imported that pass takes ~45 ms, long enough for the GPU queue to run dry (measured:
        one such pause per ~60 batches = 0.5-0.8 ms per batch).  ``gc.freeze()`` moves the objects alive now into
        the permanent generation, so later full collections only look at what was created afterwards."""
        import gc

        gc.collect()
        gc.freeze()

    def run(self, batches, depth: int = 3):
        """Generator over ``(hidden, texts[, bboxes])`` batches with ``depth`` batches in flight."""
        for item in batches:
            self.submit(*item)
            if len(self._queue) >= depth:
                yield self.result()
        while self._queue:
            yield self.result()

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

from . import decode, ops


class HeadsDecodePipeline:
    def __init__(self, decoder, device=None, score_thresh: float = 0.0, fused_spots: Optional[bool] = None):
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

    def result(self, assemble: bool = True):
        """Python results (one reference 7-tuple per document) of the oldest submitted batch.
        ``assemble=False`` returns the raw :class:`decode.DeviceDecode` (records on the host)."""
        pending, texts, bboxes = self._queue.popleft()
        t0 = time.perf_counter()
        with torch.cuda.stream(self.compute):
            dd = pending.finish()
        t1 = time.perf_counter()
        self.wait_s += t1 - t0  # host blocked on the GPU (0 when the host is the slower side)
        if not assemble:
            return dd
        out = decode.assemble_many(dd, range(dd.batch), texts, bboxes)
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

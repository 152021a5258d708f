"""Drop-in replacements for pipeline/decode.py of ZeningLin/PEneo.

``parse_matrix_spots``, ``sample_decode_peneo`` and ``decode_peneo`` keep the reference's
signatures and return objects (pipeline/decode.py:9-14, 72-85, 381-396).  Spot extraction
(softmax / argmax / max-prob / compaction) and link resolution (1-1 maps, key-value chain walk)
run as CUDA kernels through the C ABI; the only host work left is what needs Python objects:
joining token strings, merging boxes and building the ordered dicts.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .ops import COUNTERS, _TORCH_DT, _require_cuda, _stream, seq_len_from_pairs, shaking_len

NUM_HEADS = 5
# upper bound on the pairs decoded by one kernel batch of decode_peneo (8 M pairs = 64 documents of seq 512:
# ~0.5 GB of stacked fp32 logits, ~0.5 GB of int64 tags)
DECODE_GROUP_PAIRS = 8 << 20


def merge_bbox(bbox_list):
    """Union box of a list of boxes (data/data_utils.py:62-76)."""
    x0, y0, x1, y1 = list(zip(*bbox_list))
    return [min(x0), min(y0), max(x1), max(y1)]


def _thresh_f32(score_thresh: float) -> float:
    """Smallest float32 >= score_thresh, so that the device's fp32 `score < t` equals the
    reference's `float(score) < score_thresh` for every fp32 score."""
    t = np.float32(score_thresh)
    if float(t) < float(score_thresh):
        t = np.nextafter(t, np.float32(np.inf), dtype=np.float32)
    return float(t)


class DeviceDecode:
    """Result of the two decode kernels for a batch, copied to the host once."""

    def __init__(self, n: int, cap: int, batch: int, records: np.ndarray, counts: np.ndarray,
                 spots: Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]]):
        self.n, self.cap, self.batch = n, cap, batch
        self.records, self.counts, self.spots = records, counts, spots
        # documents whose spot lists overflowed `cap` were decoded again with a larger capacity: their records live
        # in `redo` (a DeviceDecode of just those documents), `redo_rows[b]` is document b's row there
        self.redo: Optional["DeviceDecode"] = None
        self.redo_rows: Dict[int, int] = {}

    def doc(self, b: int):
        if b in self.redo_rows:
            return self.redo.doc(self.redo_rows[b])
        n, cap = self.n, self.cap
        rec = self.records[b]
        hdr = rec[:16]
        off = 16
        le = rec[off : off + 2 * n].reshape(-1, 2)[: hdr[0]]
        off += 2 * n
        lgh = rec[off : off + 2 * n].reshape(-1, 2)[: hdr[1]]
        off += 2 * n
        lgt = rec[off : off + 2 * n].reshape(-1, 2)[: hdr[2]]
        off += 2 * n
        elh = rec[off : off + 2 * cap].reshape(-1, 2)[: hdr[3]]
        off += 2 * cap
        elt = rec[off : off + 2 * cap].reshape(-1, 2)[: hdr[4]]
        off += 2 * cap
        kv = rec[off : off + 4 * cap].reshape(-1, 4)[: hdr[5]]
        return le, lgh, lgt, elh, elt, kv


def _as_batched_inputs(shakings: Sequence[torch.Tensor]):
    """Normalise five [B, P, C] logits (or [B, P] / [B, P, 1] tags) to contiguous CUDA tensors
    of one dtype.  Mode follows model/peneo_decoder.py:98-104."""
    first = shakings[0]
    tag_mode = not (first.dim() > 2 and first.shape[-1] > 1)
    outs = []
    for k, t in enumerate(shakings):
        _require_cuda(t, "shaking tensor")
        if tag_mode:
            t = t.reshape(t.shape[0], -1)
            t = t if t.dtype == torch.int64 else t.long()
        else:
            if t.dtype not in (torch.float32, torch.bfloat16, torch.float16):
                t = t.float()
            if t.dtype != first.dtype:
                t = t.to(first.dtype if first.dtype in (torch.float32, torch.bfloat16, torch.float16) else torch.float32)
        outs.append(t.contiguous())
    return outs, tag_mode


class PendingDecode:
    """K3 + K4 enqueued for a batch; the compact result is on its way to pinned host memory.
    ``finish()`` waits for it; documents whose spot lists overflowed the capacity (dense predictions of an
    untrained model) are decoded again, alone, with exactly the capacity their longest list needs."""

    def __init__(self, ins, n, cap, decode_gt, score_thresh, want_spots, d2h_stream=None, k3_events=None, heads=None,
                 host_buffers=None, record_event=True, defer_d2h=False):
        # host_buffers = (counts_h, rec_h) pinned tensors supplied by the caller and record_event=False: CUDA-graph
        # capture (no pinned allocation and no event inside the captured region; the caller synchronises the replay).
        # defer_d2h: the copies of the records are not enqueued here either — the caller issues them (d2h_copy) on its
        # own stream after every replay, so that they overlap the next batch instead of sitting in the compute stream
        self.host_buffers, self.record_event, self.defer_d2h = host_buffers, record_event, defer_d2h
        # heads = (WeightPack, ab [B * n, 2d]): spot extraction fused into the pair kernel (peneo_pair_heads_spots_fwd):
        # there are no logits (`ins` is None), the compact spot lists come straight from the heads
        self.heads = heads
        self.ins, self.n, self.cap = ins, n, cap
        self.decode_gt, self.score_thresh, self.want_spots = decode_gt, score_thresh, want_spots
        self.d2h_stream = d2h_stream  # optional side stream so the record copy does not stall the next batch
        self.k3_events = k3_events    # optional list: receives (start, stop) CUDA events around the spot-extraction kernel
        self._launch()

    def _launch(self):
        lib = _lib.load()
        ins, n, cap = self.ins, self.n, self.cap
        if self.heads is not None:
            pack, ab = self.heads
            b, dev = ab.shape[0] // n, ab.device
        else:
            b, dev = ins[0].shape[0], ins[0].device
        self.batch = b
        self.spot_p = torch.empty(b * NUM_HEADS * cap, dtype=torch.int32, device=dev)
        self.spot_tag = torch.empty_like(self.spot_p)
        self.spot_score = torch.empty(b * NUM_HEADS * cap, dtype=torch.float32, device=dev)
        counts = torch.empty(b * NUM_HEADS, dtype=torch.int32, device=dev)
        if self.k3_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record(torch.cuda.current_stream(dev))
        if self.heads is not None:
            ws = torch.empty(lib.peneo_pair_heads_spots_workspace_bytes(b, n), dtype=torch.uint8, device=dev)
            _lib.check(
                lib.peneo_pair_heads_spots_fwd(pack.dims.c(), pack.prec, pack.buf.data_ptr(), ab.data_ptr(), b, n, cap,
                                               self.spot_p.data_ptr(), self.spot_tag.data_ptr(), self.spot_score.data_ptr(),
                                               counts.data_ptr(), ws.data_ptr(), _stream(dev)),
                "peneo_pair_heads_spots_fwd",
            )
            COUNTERS["kernels"] += 1  # K2 with the spots-only epilogue + the gather (counted with K4 below as K3)
        else:
            ws = torch.empty(max(16, lib.peneo_decode_spots_workspace_bytes(b, n)), dtype=torch.uint8, device=dev)
            _lib.check(
                lib.peneo_decode_spots(b, n, _lib.ptrs5(ins), _TORCH_DT[ins[0].dtype], cap, self.spot_p.data_ptr(),
                                       self.spot_tag.data_ptr(), self.spot_score.data_ptr(), counts.data_ptr(),
                                       ws.data_ptr(), _stream(dev)),
                "peneo_decode_spots",
            )
        if self.k3_events is not None:
            ev[1].record(torch.cuda.current_stream(dev))
            self.k3_events.append(ev)
        doc_ints = lib.peneo_decode_resolve_doc_ints(n, cap)
        rec = torch.empty(b, doc_ints, dtype=torch.int32, device=dev)
        ws2 = torch.empty(max(16, lib.peneo_decode_resolve_workspace_bytes(b, n)), dtype=torch.uint8, device=dev)
        _lib.check(
            lib.peneo_decode_resolve(b, n, cap, self.spot_p.data_ptr(), self.spot_tag.data_ptr(),
                                     self.spot_score.data_ptr(), counts.data_ptr(), int(self.decode_gt),
                                     _thresh_f32(self.score_thresh), rec.data_ptr(), ws2.data_ptr(), _stream(dev)),
            "peneo_decode_resolve",
        )
        COUNTERS["kernels"] += 3  # K3 spots, K4a maps, K4b links
        if self.host_buffers is not None:
            self.counts_h, self.rec_h = self.host_buffers
        else:
            self.counts_h = torch.empty(counts.shape, dtype=torch.int32, pin_memory=True)
            self.rec_h = torch.empty(rec.shape, dtype=torch.int32, pin_memory=True)
        cur = torch.cuda.current_stream(dev)
        self._dev_out = (counts, rec)
        if self.defer_d2h:
            self.event = None
        elif self.d2h_stream is not None and self.d2h_stream != cur:
            ready = torch.cuda.Event()
            ready.record(cur)
            with torch.cuda.stream(self.d2h_stream):
                self.d2h_stream.wait_event(ready)
                self.counts_h.copy_(counts, non_blocking=True)
                self.rec_h.copy_(rec, non_blocking=True)
                counts.record_stream(self.d2h_stream)
                rec.record_stream(self.d2h_stream)
                self.event = torch.cuda.Event()
                self.event.record(self.d2h_stream)
        else:
            self.counts_h.copy_(counts, non_blocking=True)
            self.rec_h.copy_(rec, non_blocking=True)
            self.event = None
            if self.record_event:
                self.event = torch.cuda.Event()
                self.event.record(cur)
        self.d2h_bytes = self.counts_h.numel() * 4 + self.rec_h.numel() * 4
        self._keep = (counts, rec, ws, ws2)  # alive until the copies have run

    def d2h_copy(self):
        """Enqueue the two record copies on the current stream (graph mode: called after every replay)."""
        counts, rec = self._dev_out
        self.counts_h.copy_(counts, non_blocking=True)
        self.rec_h.copy_(rec, non_blocking=True)

    def finish(self) -> "DeviceDecode":
        if self.event is not None:
            self.event.synchronize()
        b = self.batch
        p = shaking_len(self.n)
        redo, redo_rows = None, {}
        worst = self.counts_h.numpy().reshape(b, NUM_HEADS).max(axis=1)
        if self.cap < p and int(worst.max()) > self.cap:
            # some lists overflowed: decode those documents again (only those) with the capacity they need
            over = np.nonzero(worst > self.cap)[0]
            idx = torch.as_tensor(over, device=self.spot_p.device)
            if self.heads is not None:
                pack, ab = self.heads
                sub_ab = ab.view(b, self.n, -1).index_select(0, idx).reshape(len(over) * self.n, -1)
                sub_ins, sub_heads = None, (pack, sub_ab)
            else:
                sub_ins, sub_heads = [t.index_select(0, idx) for t in self.ins], None
            sub = PendingDecode(sub_ins, self.n, min(p, int(worst.max())), self.decode_gt, self.score_thresh,
                                self.want_spots, None, None, sub_heads)
            self.d2h_bytes += sub.d2h_bytes
            redo, redo_rows = sub.finish(), {int(d): r for r, d in enumerate(over)}
        spots = None
        if self.want_spots:
            cap = self.cap
            spots = (self.spot_p.cpu().numpy().reshape(b, NUM_HEADS, cap), self.spot_tag.cpu().numpy().reshape(b, NUM_HEADS, cap),
                     self.spot_score.cpu().numpy().reshape(b, NUM_HEADS, cap))
        if self.record_event:
            self._keep = None
        dd = DeviceDecode(self.n, self.cap, b, self.rec_h.numpy(), self.counts_h.numpy().reshape(b, NUM_HEADS), spots)
        dd.redo, dd.redo_rows = redo, redo_rows
        return dd


def device_decode_async(shakings: Sequence[torch.Tensor], n: int, decode_gt: bool = False, score_thresh: float = 0,
                        cap: Optional[int] = None, want_spots: bool = False, d2h_stream=None,
                        k3_events=None) -> PendingDecode:
    """Enqueue K3 (spots) + K4 (resolve) + the D2H copy of the compact records on the current stream."""
    ins, _tag_mode = _as_batched_inputs(shakings)
    p = shaking_len(n)
    for k, t in enumerate(ins):
        if t.shape[1] != p:
            raise ValueError(f"shaking tensor {k} has {t.shape[1]} rows, expected {p} for seq_len {n}")
    if cap is None:
        cap = min(p, max(8 * n, 1024))
    return PendingDecode(ins, n, cap, decode_gt, score_thresh, want_spots, d2h_stream, k3_events)


def heads_decode_async(pack, ab: torch.Tensor, batch: int, n: int, score_thresh: float = 0, cap: Optional[int] = None,
                       want_spots: bool = False, d2h_stream=None, k2_events=None) -> PendingDecode:
    """Pair heads + spot extraction in ONE kernel (no logits in HBM) + link resolution + the D2H copy of the compact
    records, enqueued on the current stream.  ``ab``: the per-token projections of ``ops.token_projections``.
    Same results as ``device_decode_async(ops.pair_heads(...))`` — the fused tcgen05 configuration only."""
    p = shaking_len(n)
    if ab.shape[0] != batch * n:
        raise ValueError("ab must hold batch * n token rows")
    if cap is None:
        cap = min(p, max(8 * n, 1024))
    return PendingDecode(None, n, cap, False, score_thresh, want_spots, d2h_stream, k2_events, (pack, ab))


def device_decode(shakings: Sequence[torch.Tensor], n: int, decode_gt: bool = False, score_thresh: float = 0,
                  cap: Optional[int] = None, want_spots: bool = False) -> DeviceDecode:
    """Run K3 (spots) + K4 (resolve) for a batch and bring the compact result to the host."""
    return device_decode_async(shakings, n, decode_gt, score_thresh, cap, want_spots).finish()


def _unflatten(p: np.ndarray, n: int):
    p = p.astype(np.int64)
    t = 2 * n + 1
    i = np.floor((t - np.sqrt((t * t - 8 * p).astype(np.float64))) / 2).astype(np.int64)
    i = np.clip(i, 0, n - 1)
    i = np.where(i * n - i * (i - 1) // 2 > p, i - 1, i)
    i = np.where((i + 1) * n - (i + 1) * i // 2 <= p, i + 1, i)
    return i, i + (p - (i * n - i * (i - 1) // 2))


def spots_from_device(dd: DeviceDecode, b: int, head: int) -> List[Tuple[int, int, int, float]]:
    """The reference's spot list [(i, j, tag, score)] for one document / head."""
    if b in dd.redo_rows:
        return spots_from_device(dd.redo, dd.redo_rows[b], head)
    cnt = int(dd.counts[b, head])
    sp, st, ss = dd.spots
    ii, jj = _unflatten(sp[b, head, :cnt], dd.n)
    return [(int(i), int(j), int(t), s.item()) for i, j, t, s in zip(ii, jj, st[b, head, :cnt], ss[b, head, :cnt])]


def _glue():
    """The CPython extension built next to this file (csrc/hostglue.c)."""
    try:
        from . import _hostglue
    except ImportError as e:  # pragma: no cover
        raise RuntimeError("peneo_b200._hostglue is not built: run `python -m peneo_b200.build`") from e
    return _hostglue


def _plain_bbox(bbox):
    if bbox is None:
        return None
    return bbox.tolist() if hasattr(bbox, "tolist") else bbox


def assemble_many(dd: DeviceDecode, docs: Sequence[int], texts: Sequence[List[str]], bboxes=None) -> List[Tuple]:
    """Host glue for several documents of one decoded batch: ordered dicts, line strings / boxes and
    key-value strings from the kernels' records (pipeline/decode.py:205-212, 353-378), built through
    the C API in ``_hostglue``."""
    texts = [t if isinstance(t, list) else list(t) for t in texts]
    if bboxes is not None:
        bboxes = [_plain_bbox(bx) for bx in bboxes]
        if all(bx is None for bx in bboxes):
            bboxes = None
    docs = list(docs)
    if dd.redo_rows and any(d in dd.redo_rows for d in docs):
        # documents decoded a second time (capacity overflow) come from dd.redo; keep the caller's order
        out = [None] * len(docs)
        for src, rows in ((dd.redo, dd.redo_rows), (dd, None)):
            pos = [k for k, d in enumerate(docs) if (d in dd.redo_rows) == (rows is not None)]
            if not pos:
                continue
            part = assemble_many(src if rows is not None else _without_redo(dd),
                                 [rows[docs[k]] if rows is not None else docs[k] for k in pos],
                                 [texts[k] for k in pos], None if bboxes is None else [bboxes[k] for k in pos])
            for k, r in zip(pos, part):
                out[k] = r
        return out
    rec = dd.records
    return _glue().assemble(memoryview(rec).cast("B"), rec.shape[1], dd.n, dd.cap, docs, texts, bboxes)


def _without_redo(dd: DeviceDecode) -> DeviceDecode:
    plain = DeviceDecode(dd.n, dd.cap, dd.batch, dd.records, dd.counts, dd.spots)
    return plain


def _assemble(dd: DeviceDecode, b: int, text: List[str], bbox):
    return assemble_many(dd, [b], [text], None if bbox is None else [bbox])[0]


def parse_matrix_spots(matrix_spots, top_score_only: bool = False, triu_mode: bool = False, score_thresh: float = 0):
    """pipeline/decode.py:9-69 for callers that hold a Python spot list (tiny, host-side by nature:
    its input and output are Python objects).  The batched decode path does this step on the GPU."""
    spot_map: dict = {}
    for head_idx, tail_idx, tag, score in matrix_spots:
        if tag == 0 or score < score_thresh:
            continue
        if triu_mode and tag == 2:
            head_idx, tail_idx = tail_idx, head_idx
        if not top_score_only:
            spot_map.setdefault(head_idx, []).append(tail_idx)
        elif head_idx not in spot_map or score > spot_map[head_idx][1]:
            spot_map[head_idx] = (tail_idx, score)
    if not top_score_only:
        return spot_map
    reverse: dict = {}
    for k, (v, s) in spot_map.items():
        if v not in reverse or s > reverse[v][1]:
            reverse[v] = (k, s)
    return {k: v for v, (k, _s) in reverse.items()}


def _infer_seq_len(seq_len, shaking_ind2matrix_ind, shaking: torch.Tensor) -> int:
    if shaking_ind2matrix_ind is not None:
        return seq_len_from_pairs(len(shaking_ind2matrix_ind))
    if seq_len is not None:
        return int(seq_len)
    raise AssertionError("seq_len or shaking_ind2matrix_ind must be given")  # pipeline/decode.py:145-147


def _to_device(t, device):
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    return t if t.is_cuda else t.to(device)


def sample_decode_peneo(
    handshaking_tagger,
    text: List[str],
    line_extraction_shaking: torch.Tensor,
    ent_linking_h2h_shaking: torch.Tensor,
    ent_linking_t2t_shaking: torch.Tensor,
    line_grouping_h2h_shaking: torch.Tensor,
    line_grouping_t2t_shaking: torch.Tensor,
    bbox: torch.Tensor = None,
    seq_len: int = None,
    shaking_ind2matrix_ind: List[Tuple[int]] = None,
    decode_gt: bool = False,
    score_thresh: float = 0,
) -> Tuple:
    """Decode one sample (pipeline/decode.py:72-378).  Returns the reference's 7-tuple
    (kv_pairs, lines, LE dict, EL-head dict, EL-tail dict, LG-head dict, LG-tail dict)."""
    sh = [line_extraction_shaking, ent_linking_h2h_shaking, ent_linking_t2t_shaking, line_grouping_h2h_shaking,
          line_grouping_t2t_shaking]
    n = _infer_seq_len(seq_len, shaking_ind2matrix_ind, sh[0])
    device = next((t.device for t in sh if torch.is_tensor(t) and t.is_cuda), torch.device("cuda"))
    sh = [_to_device(t, device).unsqueeze(0) for t in sh]
    dd = device_decode(sh, n, decode_gt=decode_gt, score_thresh=score_thresh)
    return _assemble(dd, 0, text, bbox)


def decode_peneo(
    handshaking_tagger,
    texts: List[List[str]],
    line_extraction_shaking_outputs,
    ent_linking_h2h_shaking_outputs,
    ent_linking_t2t_shaking_outputs,
    line_grouping_h2h_shaking_outputs,
    line_grouping_t2t_shaking_outputs,
    line_extraction_shaking_tags,
    ent_linking_h2h_shaking_tags,
    ent_linking_t2t_shaking_tags,
    line_grouping_h2h_shaking_tags,
    line_grouping_t2t_shaking_tags,
    orig_bboxes,
    file_ids: List[str],
):
    """Decode predictions and ground truth of a list of samples (pipeline/decode.py:381-511).
    Samples of equal length are decoded together in one pair of kernel launches."""
    outs = [line_extraction_shaking_outputs, ent_linking_h2h_shaking_outputs, ent_linking_t2t_shaking_outputs,
            line_grouping_h2h_shaking_outputs, line_grouping_t2t_shaking_outputs]
    tags = [line_extraction_shaking_tags, ent_linking_h2h_shaking_tags, ent_linking_t2t_shaking_tags,
            line_grouping_h2h_shaking_tags, line_grouping_t2t_shaking_tags]
    count = min(len(file_ids), len(texts), len(orig_bboxes), *[len(o) for o in outs], *[len(t) for t in tags])
    all_pred, all_gt, all_ids = [None] * count, [None] * count, []
    if len(texts) == 0:  # sic: the reference tests len(texts) (pipeline/decode.py:471)
        return [], [], []
    by_len: Dict[int, List[int]] = {}
    for s in range(count):
        by_len.setdefault(len(orig_bboxes[s]), []).append(s)
    device = next((o[0].device for o in outs if len(o) and torch.is_tensor(o[0]) and o[0].is_cuda), torch.device("cuda"))
    for n, group in by_len.items():
        # Bounded kernel batches: stacking copies every logit / tag tensor of the group once more on the device, and
        # the compact buffers scale with batch * capacity, so a group is decoded in slices of at most
        # DECODE_GROUP_PAIRS pairs (the reference decodes sample by sample; the kernels' grid also caps a batch
        # at 65 535 documents).
        step = max(1, min(65535, DECODE_GROUP_PAIRS // max(1, shaking_len(n))))
        for lo in range(0, len(group), step):
            idxs = group[lo : lo + step]
            pred_in = [torch.stack([_to_device(o[s], device) for s in idxs]) for o in outs]
            dp = device_decode(pred_in, n, decode_gt=False)
            del pred_in
            gt_in = [torch.stack([_to_device(t[s], device) for s in idxs]) for t in tags]
            dg = device_decode(gt_in, n, decode_gt=True)
            del gt_in
            rows = list(range(len(idxs)))
            tx = [texts[s] for s in idxs]
            for s, pr, gt in zip(idxs, assemble_many(dp, rows, tx), assemble_many(dg, rows, tx)):
                all_pred[s], all_gt[s] = pr, gt
    all_ids = [file_ids[s] for s in range(count)]
    return all_pred, all_gt, all_ids

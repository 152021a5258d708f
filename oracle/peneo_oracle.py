"""TEST INFRASTRUCTURE ONLY — CPU oracle for the PEneo hot path (heads, loss, decode).

This file restates, on the CPU, the algorithm of ZeningLin/PEneo's decoder so that the CUDA
path in ``peneo_b200/`` can be checked against it.  It is *not* part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``peneo_b200``) never imports anything from ``oracle/``
and raises if its CUDA library is missing.

Parity pin: the reference ships no tests or golden vectors, so the oracle is
pinned against *outputs of the reference itself*: ``oracle/make_golden.py`` imports the real
reference (``oracle/ref_shim.py``) in the build container and writes ``tests/golden/*.pt``;
``tests/test_oracle_golden.py`` checks every function here against those fixtures (and against
the live reference when ``/root/reference`` is present).

Reference lines followed (paths relative to the reference repo):
  * shrink projection ............ model/peneo_decoder.py:213-222, 349-350
  * handshaking kernel ........... model/peneo_decoder.py:129-177
  * classifier heads ............. model/peneo_decoder.py:231-292, 355-363
  * loss (OHEM off / on) ......... model/custom_loss.py:189-288, model/peneo_decoder.py:315-336, 375-428
  * tag codec .................... model/peneo_decoder.py:34-115
  * decode ....................... pipeline/decode.py:9-69, 72-378, 381-511
  * merge_bbox ................... data/data_utils.py:62-76
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

HEAD_NAMES = (
    "line_extraction",
    "ent_linking_h2h",
    "ent_linking_t2t",
    "line_grouping_h2h",
    "line_grouping_t2t",
)
HEAD_CLASSES = (2, 3, 3, 3, 3)


# --------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------
def split_params(state_dict: Dict[str, torch.Tensor], dtype=torch.float64) -> dict:
    """Pick the decoder parameters out of a reference-keyed state dict.

    Key names are the checkpoint interface listed in SURVEY.md §5 (model/peneo_decoder.py:213-292).
    ``num_layers`` is inferred from the keys: ``<head>_fc.weight`` (L == 1) or
    ``<head>_fc.{0,3,..}.weight`` (L >= 2).
    """
    return _split_live({k: v.detach().to("cpu", dtype) for k, v in state_dict.items()})


def silu(x: torch.Tensor) -> torch.Tensor:
    return x * torch.sigmoid(x)


def shaking_len(n: int) -> int:
    return n * (n + 1) // 2


def shaking_index(i: int, j: int, n: int) -> int:
    """Row-major upper-triangular flat index of (i <= j); SURVEY.md appendix A.3."""
    return i * n - i * (i - 1) // 2 + (j - i)


# --------------------------------------------------------------------------------------
# dropout masks (training mode).  torch's RNG stream cannot be matched by a CUDA kernel, so the product
# defines its own counter-based mask (peneo_b200/csrc/common.cuh); this is its restatement, so that the
# training-mode forward / backward can be checked element for element with the SAME mask.
# --------------------------------------------------------------------------------------
SITE_TOK0, SITE_TOK1 = 0, 1


def site_head(head: int, layer: int) -> int:
    return 16 + 8 * head + layer


def _mix32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64) & 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def dropout_mask(dropout, site: int, rows: np.ndarray, ncols: int, dtype=torch.float64) -> torch.Tensor:
    """[len(rows), ncols] tensor of 0 / (1 / (1 - p)) for the Dropout module `site`; `rows` are token indices
    (per-token sites) or batch-flat pair indices (pair sites).  ``dropout = (p, seed)``."""
    p, seed = dropout
    p32 = np.float32(p)
    thresh = min(int(float(p32) * 4294967296.0), 0xFFFFFFFF)
    scale = float(np.float32(1.0) / (np.float32(1.0) - p32)) if p32 < 1 else 0.0
    seed_lo, seed_hi = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    inner = _mix32(np.array([(seed_hi + 0x9E3779B9 * (site + 1)) & 0xFFFFFFFF], dtype=np.uint64))
    key = _mix32(np.array([seed_lo], dtype=np.uint64) ^ inner)[0]
    r = (np.asarray(rows, dtype=np.uint64)[:, None] * 0x9E3779B1 + np.arange(ncols, dtype=np.uint64)[None, :]) & 0xFFFFFFFF
    keep = _mix32(r ^ key) >= thresh
    return torch.from_numpy(keep.astype(np.float64) * scale).to(dtype)


# --------------------------------------------------------------------------------------
# heads
# --------------------------------------------------------------------------------------
def shrink_projection(p: dict, x: torch.Tensor, dropout=None) -> torch.Tensor:
    """model/peneo_decoder.py:215-222; eval mode when ``dropout`` is None, else Dropout after each SiLU with the
    product's mask (token index = position in the batch-flattened [B * N] token list)."""
    if p["shrink"] is None:
        return x
    w1, b1, w2, b2 = p["shrink"]
    y = silu(x @ w1.T + b1)
    if dropout is not None:
        rows = np.arange(x.shape[0] * x.shape[1])
        y = y * dropout_mask(dropout, SITE_TOK0, rows, y.shape[-1], y.dtype).reshape(y.shape)
    y = silu(y @ w2.T + b2)
    if dropout is not None:
        y = y * dropout_mask(dropout, SITE_TOK1, rows, y.shape[-1], y.dtype).reshape(y.shape)
    return y


def handshake_ref_style(p: dict, y: torch.Tensor) -> torch.Tensor:
    """model/peneo_decoder.py:149-177 op for op: materialise all ordered pairs, keep i <= j
    in row-major order, apply combine_fc + SiLU.  O(N^2 * 2D) memory — small N only."""
    b, n, d = y.shape
    rows = y.unsqueeze(2).expand(b, n, n, d)  # [b, i, j, :] = y_i
    cols = y.unsqueeze(1).expand(b, n, n, d)  # [b, i, j, :] = y_j
    pairs = torch.cat([rows, cols], dim=-1)  # first D columns <- row token i
    iu = torch.triu_indices(n, n)  # row-major (i <= j), same order as nonzero(mask)
    pairs = pairs[:, iu[0], iu[1], :]
    return silu(pairs @ p["combine_w"].T + p["combine_b"])


def classifier(layers, s: torch.Tensor, dropout=None, head: int = 0, rows=None) -> torch.Tensor:
    """model/peneo_decoder.py:253-271; eval mode when ``dropout`` is None, else Dropout after every hidden SiLU
    (``rows`` = batch-flat pair index of every row of ``s`` flattened to 2-D)."""
    h = s
    for li, (w, b) in enumerate(layers):
        h = h @ w.T + b
        if li + 1 < len(layers):
            h = silu(h)
            if dropout is not None:
                h = h * dropout_mask(dropout, site_head(head, li), rows, h.shape[-1], h.dtype).reshape(h.shape)
    return h


def heads_ref_style(p: dict, x: torch.Tensor, dropout=None) -> List[torch.Tensor]:
    """Full forward in the reference's own op order (model/peneo_decoder.py:349-363).
    Returns the 5 logits tensors in return order LE, EL-h2h, EL-t2t, LG-h2h, LG-t2t."""
    y = shrink_projection(p, x, dropout)
    s = handshake_ref_style(p, y)
    rows = np.arange(s.shape[0] * s.shape[1]) if dropout is not None else None
    return [classifier(layers, s, dropout, k, rows) for k, layers in enumerate(p["heads"])]


def token_projections(p: dict, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """A_i = W_c[:, :D] y_i ;  Bm_j = W_c[:, D:] y_j + b_c  (SURVEY.md §8a-H3, appendix B)."""
    y = shrink_projection(p, x)
    d = y.shape[-1]
    a = y @ p["combine_w"][:, :d].T
    bm = y @ p["combine_w"][:, d:].T + p["combine_b"]
    return a, bm


def heads_chunked(p: dict, x: torch.Tensor, row_block: int = 16) -> List[torch.Tensor]:
    """Same function as :func:`heads_ref_style` through the additive decomposition
    ``S[p(i,j)] = SiLU(A_i + Bm_j)`` evaluated a block of rows at a time, so memory is
    O(row_block * N * D).  Used for N >= 1024 and as the fp64 yardstick."""
    a, bm = token_projections(p, x)
    b, n, d = a.shape
    outs = [
        torch.empty(b, shaking_len(n), layers[-1][0].shape[0], dtype=x.dtype) for layers in p["heads"]
    ]
    for i0 in range(0, n, row_block):
        i1 = min(n, i0 + row_block)
        for i in range(i0, i1):
            s = silu(a[:, i : i + 1, :] + bm[:, i:, :])  # pairs (i, i..n-1)
            p0 = shaking_index(i, i, n)
            for k, layers in enumerate(p["heads"]):
                outs[k][:, p0 : p0 + (n - i), :] = classifier(layers, s)
    return outs


# --------------------------------------------------------------------------------------
# loss
# --------------------------------------------------------------------------------------
def weighted_nll(logits: torch.Tensor, target: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """Per-element ``w[t] * (logsumexp(x) - x[t])`` in fp32-or-better (custom_loss.py:204-210)."""
    x = logits.reshape(-1, logits.shape[-1])
    x = x.float() if x.dtype in (torch.float16, torch.bfloat16) else x
    t = target.reshape(-1)
    lse = torch.logsumexp(x, dim=-1)
    picked = x.gather(1, t.unsqueeze(1)).squeeze(1)
    return weight.to(x.dtype)[t] * (lse - picked)


def ce_loss(
    logits: torch.Tensor,
    target: torch.Tensor,
    weight: torch.Tensor,
    num_hard_positive: int = -1,
    num_hard_negative: int = -1,
) -> torch.Tensor:
    """CrossEntropyLossOHEM.forward with reduction='mean' (custom_loss.py:189-288).

    OHEM off (both -1): weighted mean over *all* flattened elements, no padding mask.
    OHEM on: reproduces the reference's observable quirks:
      (i) the kept set is ``sorted_loss[sorted_index[:k]]`` — sorted values indexed with
          unsorted-array positions — not the top-k;
      (ii) a side with k <= 0 keeps everything, yet the divisor is ``k_pos + k_neg`` with the
          raw (possibly -1 -> min(n,-1) = -1) values.
    """
    per = weighted_nll(logits, target, weight)
    t = target.reshape(-1)
    if num_hard_positive == -1 and num_hard_negative == -1:
        return per.sum() / weight.to(per.dtype)[t].sum()

    neg_mask = t == 0
    pos, neg = per[~neg_mask], per[neg_mask]

    def keep(side: torch.Tensor, k_cfg: int):
        srt, idx = torch.sort(side, descending=True)
        k = min(srt.shape[0], k_cfg)
        if 0 < k < srt.shape[0]:
            srt = srt[idx[:k]]  # quirk (i)
        return srt, k

    pos_kept, kp = keep(pos, num_hard_positive)
    neg_kept, kn = keep(neg, num_hard_negative)
    return (pos_kept.sum() + neg_kept.sum()) / (kp + kn)


def decoder_loss(
    logits: Sequence[torch.Tensor],
    tags: Sequence[torch.Tensor],
    category_weights: Sequence[float],
    loss_ratio: Optional[Sequence[float]] = None,
    num_hard_positive: int = -1,
    num_hard_negative: int = -1,
) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """model/peneo_decoder.py:375-428: five sub-losses (LE uses weights[:-1]) and their
    ratio-weighted sum.  ``logits`` / ``tags`` are in return order (LE, ELh, ELt, LGh, LGt)."""
    dt = logits[0].dtype if logits[0].dtype == torch.float64 else torch.float32
    w_link = torch.tensor(list(category_weights), dtype=dt)
    w_le = w_link[:-1]
    ratios = [1.0] * 5 if loss_ratio is None else list(loss_ratio)
    subs = []
    for k in range(5):
        w = w_le if k == 0 else w_link
        subs.append(ce_loss(logits[k], tags[k], w, num_hard_positive, num_hard_negative))
    total = sum(r * s for r, s in zip(ratios, subs))
    return total, subs


def loss_and_grads(
    state_dict: Dict[str, torch.Tensor],
    x: torch.Tensor,
    tags: Sequence[torch.Tensor],
    category_weights: Sequence[float],
    loss_ratio: Optional[Sequence[float]] = None,
    dtype=torch.float64,
    dropout=None,
):
    """Backward oracle: autograd through the restated forward, fp64 by default (``dropout = (p, seed)`` applies
    the product's training-mode masks; None = dropout off).
    Returns (loss, sub_losses, {param_key: grad}, d_sequence_output)."""
    sd = {}
    for k, v in state_dict.items():
        t = v.detach().to("cpu", dtype).clone()
        if "_loss." not in k:  # link_loss.weight / le_loss.weight are buffers
            t.requires_grad_(True)
        sd[k] = t
    p = _split_live(sd)
    xx = x.detach().to("cpu", dtype).requires_grad_(True)
    if dropout is not None:
        logits = heads_ref_style(p, xx, dropout)
    else:
        logits = heads_chunked(p, xx, row_block=64) if xx.shape[1] > 64 else heads_ref_style(p, xx)
    total, subs = decoder_loss(logits, tags, category_weights, loss_ratio)
    total.backward()
    grads = {k: v.grad for k, v in sd.items() if v.requires_grad and v.grad is not None}
    return total.detach(), [s.detach() for s in subs], grads, xx.grad


def _split_live(sd: Dict[str, torch.Tensor]) -> dict:
    p = {"shrink": None, "heads": []}
    if "shrink_projection.0.weight" in sd:
        p["shrink"] = (
            sd["shrink_projection.0.weight"],
            sd["shrink_projection.0.bias"],
            sd["shrink_projection.3.weight"],
            sd["shrink_projection.3.bias"],
        )
    p["combine_w"] = sd["handshaking_kernel.combine_fc.weight"]
    p["combine_b"] = sd["handshaking_kernel.combine_fc.bias"]
    for name in HEAD_NAMES:
        layers = []
        if f"{name}_fc.weight" in sd:
            layers.append((sd[f"{name}_fc.weight"], sd[f"{name}_fc.bias"]))
        else:
            idx = 0
            while f"{name}_fc.{idx}.weight" in sd:
                layers.append((sd[f"{name}_fc.{idx}.weight"], sd[f"{name}_fc.{idx}.bias"]))
                idx += 3  # Linear, SiLU, Dropout triples (model/peneo_decoder.py:258-269)
        p["heads"].append(layers)
    return p


# --------------------------------------------------------------------------------------
# tag codec
# --------------------------------------------------------------------------------------
def spots_to_tags(batch_spots: Sequence[Sequence[Tuple[int, int, int]]], seq_len: int) -> torch.Tensor:
    """HandshakingTaggingScheme.spots2shaking_tag4batch (model/peneo_decoder.py:34-73):
    dense [B, P] int64 tags; later spots overwrite earlier ones at the same cell."""
    out = torch.zeros(len(batch_spots), shaking_len(seq_len), dtype=torch.int64)
    for b, spots in enumerate(batch_spots):
        for sp in spots:
            # the reference looks (i, j) up in an N x N list-of-lists that is 0 below the diagonal (lines 55-59, 70):
            # Python indexing, so indices in [-N, N) are legal and anything else raises IndexError
            i, j = sp[0], sp[1]
            if not (-seq_len <= i < seq_len and -seq_len <= j < seq_len):
                raise IndexError("list index out of range")
            i, j = i % seq_len, j % seq_len
            out[b, shaking_index(i, j, seq_len) if i <= j else 0] = sp[2]
    return out


def unflatten_index(pidx: np.ndarray, n: int) -> Tuple[np.ndarray, np.ndarray]:
    """Inverse of :func:`shaking_index` (vectorised, exact integer arithmetic)."""
    pidx = np.asarray(pidx, dtype=np.int64)
    # largest i with i*n - i*(i-1)/2 <= p
    t = 2 * n + 1
    i = np.floor((t - np.sqrt((t * t - 8 * pidx).astype(np.float64))) / 2).astype(np.int64)
    i = np.clip(i, 0, n - 1)
    start = i * n - i * (i - 1) // 2
    i = np.where(start > pidx, i - 1, i)
    start = i * n - i * (i - 1) // 2
    nxt = (i + 1) * n - (i + 1) * i // 2
    i = np.where(nxt <= pidx, i + 1, i)
    start = i * n - i * (i - 1) // 2
    return i, i + (pidx - start)


def get_spots(shaking: torch.Tensor, seq_len: int) -> List[Tuple[int, int, int, float]]:
    """HandshakingTaggingScheme.get_spots_from_shaking_tag (model/peneo_decoder.py:75-115).

    2-D input with a class dimension > 1: softmax in the input's dtype, argmax *of the softmax*
    (ties -> lowest class), score = max prob.  Otherwise (GT tags): pred = tag, score = 1.
    Spots are emitted for every non-zero pred in increasing flat index."""
    if shaking.dim() > 1 and shaking.shape[-1] > 1:
        prob = shaking.softmax(-1)
        pred = prob.argmax(-1)
        score = prob.max(dim=-1)[0]
    else:
        pred = shaking.reshape(-1) if shaking.dim() > 1 else shaking
        score = torch.ones_like(pred)
    nz = torch.nonzero(pred).reshape(-1).numpy()
    ii, jj = unflatten_index(nz, seq_len)
    pred_np = pred.numpy()
    score_np = score.float().numpy() if score.dtype in (torch.bfloat16, torch.float16) else score.numpy()
    return [
        (int(i), int(j), int(pred_np[q]), score_np[q].item()) for q, i, j in zip(nz, ii, jj)
    ]


# --------------------------------------------------------------------------------------
# decode
# --------------------------------------------------------------------------------------
def merge_bbox(bbox_list):
    """data/data_utils.py:62-76."""
    x0, y0, x1, y1 = list(zip(*bbox_list))
    return [min(x0), min(y0), max(x1), max(y1)]


def parse_matrix_spots(spots, top_score_only=False, triu_mode=False, score_thresh=0):
    """pipeline/decode.py:9-69.  Ordered-dict semantics are part of the contract:
    stage 1 keeps, per head, the strictly best-scoring tail (first arrival wins ties);
    stage 2 keeps, per tail, the strictly best-scoring head, visiting stage-1 entries in
    insertion order; the result is keyed by head in stage-2 insertion order.  A head that
    loses its tail in stage 2 is dropped (no fallback)."""
    first: Dict[int, object] = {}
    for head, tail, tag, score in spots:
        if tag == 0 or score < score_thresh:
            continue
        if triu_mode and tag == 2:
            head, tail = tail, head
        if not top_score_only:
            first.setdefault(head, []).append(tail)
        elif head not in first or score > first[head][1]:
            first[head] = (tail, score)
    if not top_score_only:
        return first
    by_tail: Dict[int, Tuple[int, float]] = {}
    for head, (tail, score) in first.items():
        if tail not in by_tail or score > by_tail[tail][1]:
            by_tail[tail] = (head, score)
    return {head: tail for tail, (head, _s) in by_tail.items()}


def _walk(head, tail, le, lg_head, lg_tail):
    """Chain walk of pipeline/decode.py:248-296 (identical for key and value).  Returns the list
    of (head, tail) line segments and the final tail.  Cycles of length >= 2 are not detected:
    the walk runs until ``num_op > 1000``."""
    segs = [(head, tail)]
    cur_h, cur_t = head, tail
    nxt = lg_head.get(cur_h, None)
    ops = 0
    while nxt is not None:
        ops += 1
        if ops > 1000 or nxt == cur_h:
            break
        nxt_t = le.get(nxt, None)
        if nxt_t is None or lg_tail.get(cur_t, None) != nxt_t:
            break
        segs.append((nxt, nxt_t))
        cur_h, cur_t = nxt, nxt_t
        nxt = lg_head.get(cur_h)
    return segs, cur_t


def sample_decode(
    text: List[str],
    shakings: Sequence[torch.Tensor],
    seq_len: int,
    bbox=None,
    decode_gt: bool = False,
    score_thresh: float = 0,
):
    """pipeline/decode.py:72-378 for one sample.  ``shakings`` in return order
    (LE, EL-h2h, EL-t2t, LG-h2h, LG-t2t).  Returns the reference's 7-tuple."""
    le_sp, elh_sp, elt_sp, lgh_sp, lgt_sp = [get_spots(s, seq_len) for s in shakings]
    top = not decode_gt
    le = parse_matrix_spots(le_sp, top, False, score_thresh)
    lg_tail = parse_matrix_spots(lgt_sp, top, True, score_thresh)
    lg_head = parse_matrix_spots(lgh_sp, top, True, score_thresh)
    if decode_gt:
        le = {k: v[0] for k, v in le.items()}
        lg_tail = {k: v[0] for k, v in lg_tail.items()}
        lg_head = {k: v[0] for k, v in lg_head.items()}
    if bbox is not None:
        bbox = bbox.tolist() if hasattr(bbox, "tolist") else bbox

    def seg_text(h, t):
        return "".join(text[h : t + 1])

    lines = []
    for h, t in le.items():
        lines.append((seg_text(h, t), merge_bbox(bbox[h : t + 1])) if bbox is not None else seg_text(h, t))

    el_tail = parse_matrix_spots(elt_sp, False, True, score_thresh)
    el_head: Dict[int, List[int]] = {}
    pairs = []
    for kh, vh, tag, score in elh_sp:
        if tag == 0 or score < score_thresh:
            continue
        if tag == 2:
            kh, vh = vh, kh
        el_head.setdefault(kh, []).append(vh)
        kt, vt = le.get(kh, None), le.get(vh, None)
        if kt is None or vt is None:
            continue
        ksegs, k_last = _walk(kh, kt, le, lg_head, lg_tail)
        vsegs, v_last = _walk(vh, vt, le, lg_head, lg_tail)
        ok_tails = el_tail.get(k_last, None)
        if ok_tails is not None and v_last in ok_tails:
            ktxt = "".join(seg_text(h, t) for h, t in ksegs).strip()
            vtxt = "".join(seg_text(h, t) for h, t in vsegs).strip()
            if bbox is not None:
                kbox = merge_bbox([merge_bbox(bbox[h : t + 1]) for h, t in ksegs])
                vbox = merge_bbox([merge_bbox(bbox[h : t + 1]) for h, t in vsegs])
                pairs.append((ktxt, vtxt, kbox, vbox))
            else:
                pairs.append((ktxt, vtxt))
    return pairs, lines, le, el_head, el_tail, lg_head, lg_tail


def decode_batch(texts, outputs, tags, orig_bboxes, file_ids):
    """pipeline/decode.py:381-511: per-sample prediction decode + GT decode.
    ``outputs`` / ``tags``: 5-sequences of per-sample tensors, return order."""
    preds, gts, ids = [], [], []
    for s in range(len(file_ids)):
        if len(texts) == 0:  # sic: the reference tests len(texts), decode.py:471
            continue
        n = len(orig_bboxes[s])
        preds.append(sample_decode(texts[s], [o[s] for o in outputs], n, decode_gt=False))
        gts.append(sample_decode(texts[s], [t[s] for t in tags], n, decode_gt=True))
        ids.append(file_ids[s])
    return preds, gts, ids


# --------------------------------------------------------------------------------------
# algorithmic work — shared by tests and bench bookkeeping
# --------------------------------------------------------------------------------------
def heads_flops(n: int, hin: int = 768, d: int = 384, hid: int = 768) -> float:
    p = shaking_len(n)
    return 2.0 * n * (hin * hid + hid * d + 2 * d * d) + 10.0 * p * d * d + 28.0 * p * d

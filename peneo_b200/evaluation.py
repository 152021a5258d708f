"""Drop-in replacements for pipeline/evaluation.py of ZeningLin/PEneo ("next" row N2 of the scope table).

``calculate_KVPE_metric`` and ``calculate_detail_KVPE_metric`` keep the reference's signatures, return
dictionaries, key order, float-typed counters and detail records (pipeline/evaluation.py:98-207, 210-665).
Two things change underneath:

* matching: the reference tests ``pred_item in gt_list`` for every prediction (O(pred x gt) per file,
  pipeline/evaluation.py:31-33, 71-80); here the ground truth is hashed once (O(pred + gt)), falling back to
  the list scan for unhashable items (pairs carrying box lists);
* cross-rank gather: instead of pickling Python lists through ``all_gather_object``
  (pipeline/evaluation.py:150-156, 416-422) every rank contributes one fixed-width int64 row per file
  (a 64-bit digest of the file name + the counters) through a tensor ``all_gather``; duplicates created by
  the distributed sampler are dropped by file name exactly as in the reference (first occurrence in rank
  order wins, pipeline/evaluation.py:173-175, 483-485).
"""
from __future__ import annotations

import hashlib
from typing import Dict, List, Sequence, Tuple, Union

import torch
import torch.distributed as dist


def _membership(gt: Sequence):
    """``x in gt`` as a callable, hashed when the items allow it."""
    try:
        s = set(gt)
    except TypeError:
        return lambda x: x in gt
    def test(x):
        try:
            return x in s
        except TypeError:  # an unhashable prediction can still equal no hashable ground truth item... or one
            return x in gt
    return test


def _prf(num_correct: float, num_pred: float, num_gt: float):
    precision = num_correct / num_pred if num_pred > 0 else 0.0
    recall = num_correct / num_gt if num_gt > 0 else 0.0
    f1 = (2 * precision * recall) / (precision + recall) if precision + recall > 0 else 0.0
    return precision, recall, f1


def _calculate_linking_metric_core(pred: Union[Dict, List], gt: Union[Dict, List]):
    """pipeline/evaluation.py:6-42."""
    if isinstance(pred, dict):
        pred = [(k, v) for k, v in pred.items()]
    if isinstance(gt, dict):
        gt = [(k, v) for k, v in gt.items()]
    num_pred, num_gt = 0.0 + len(pred), 0.0 + len(gt)
    inside = _membership(gt)
    num_correct = 0.0
    for item in pred:
        if inside(item):
            num_correct += 1
    precision, recall, f1 = _prf(num_correct, num_pred, num_gt)
    return precision, recall, f1, num_pred, num_gt, num_correct


def _calculate_KV_metric_core(pred: List, gt: List, return_detail: bool = False):
    """pipeline/evaluation.py:45-95 (TP / FP records in prediction order, then FN records in GT order)."""
    num_pred, num_gt, num_correct = 0.0 + len(pred), 0.0 + len(gt), 0.0
    detail, matched = [], []
    inside = _membership(gt)
    for p in pred:
        if inside(p):
            num_correct += 1
            if return_detail:
                detail.append({"status": "TP", "pred": p})
            matched.append(p)
        elif return_detail:
            detail.append({"status": "FP", "pred": p})
    precision, recall, f1 = _prf(num_correct, len(pred), len(gt))
    if return_detail:
        was_matched = _membership(matched)
        for g in gt:
            if not was_matched(g):
                detail.append({"status": "FN", "gt": g})
        return precision, recall, f1, num_pred, num_gt, num_correct, detail
    return precision, recall, f1, num_pred, num_gt, num_correct


def _digest(fname) -> int:
    h = hashlib.blake2b(str(fname).encode("utf-8", "surrogatepass"), digest_size=8).digest()
    return int.from_bytes(h, "little", signed=True)


def _gather_rows(rows: List[List], width: int) -> List[List]:
    """All ranks' ``[fname, c_1 .. c_width]`` rows, concatenated in rank order.  Single process: identity.
    Distributed: one int64 tensor all_gather of ``[digest, counters...]`` rows (padded to the longest rank)."""
    if not (dist.is_available() and dist.is_initialized()):
        return rows
    dist.barrier()
    world = dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    n_local = torch.tensor([len(rows)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    n_max = max(int(c.item()) for c in counts)
    buf = torch.zeros(max(n_max, 1), 1 + width, dtype=torch.int64, device=dev)
    if rows:
        buf[: len(rows)] = torch.tensor([[_digest(r[0])] + [int(v) for v in r[1:]] for r in rows], dtype=torch.int64)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out: List[List] = []
    for part, c in zip(parts, counts):
        for r in part[: int(c.item())].tolist():
            out.append([("digest", r[0])] + [float(v) for v in r[1:]])
    return out


def _accumulate(rows: List[List], width: int):
    seen, totals, processed = set(), [0.0] * width, 0
    for row in rows:
        key = row[0]
        if key in seen:  # avoid duplication with the distributed sampler
            continue
        seen.add(key)
        for q in range(width):
            totals[q] += row[1 + q]
        processed += 1
    return totals, processed


def calculate_KVPE_metric(all_pred: List[Tuple], all_gt: List[Tuple], all_fname: List[str]):
    """pipeline/evaluation.py:98-207."""
    sample_detail, rows = [], []
    for fname, pred, gt in zip(all_fname, all_pred, all_gt):
        p, r, f1, n_pred, n_gt, n_correct, info = _calculate_KV_metric_core(pred[0], gt[0], return_detail=True)
        sample_detail.append({"fname": fname, "num_pred": n_pred, "num_gt": n_gt, "num_correct": n_correct,
                              "precision": p, "recall": r, "f1": f1, "detail": info})
        rows.append([fname, n_pred, n_gt, n_correct])
    (num_pred, num_gt, num_correct), processed = _accumulate(_gather_rows(rows, 3), 3)
    precision, recall, f1 = _prf(num_correct, num_pred, num_gt)
    detail = {"precision": precision, "recall": recall, "f1": f1, "num_pred": num_pred, "num_gt": num_gt,
              "num_correct": num_correct, "num_sample_processed": processed, "detail": sample_detail}
    return {"precision": precision, "recall": recall, "f1": f1}, detail


_TASKS = ("kv_pair", "line_extraction", "ent_linking_head", "ent_linking_tail", "line_grouping_head",
          "line_grouping_tail")


def _flatten_multimap(d: Dict) -> List[Tuple]:
    return [(k, v) for k, vs in d.items() for v in vs]


def calculate_detail_KVPE_metric(all_pred: List[Tuple], all_gt: List[Tuple], all_fname: List[str]):
    """pipeline/evaluation.py:210-665: key-value pairs, line extraction, entity linking head / tail and line
    grouping head / tail."""
    sample_details, rows = [], []
    for fname, pred, gt in zip(all_fname, all_pred, all_gt):
        p_kv, p_lines, _, p_elh, p_elt, p_lgh, p_lgt = pred
        g_kv, g_lines, _, g_elh, g_elt, g_lgh, g_lgt = gt
        kv = _calculate_KV_metric_core(p_kv, g_kv, return_detail=True)
        res = [
            kv[:6],
            _calculate_KV_metric_core(p_lines, g_lines, return_detail=False),
            _calculate_linking_metric_core(_flatten_multimap(p_elh), _flatten_multimap(g_elh)),
            _calculate_linking_metric_core(_flatten_multimap(p_elt), _flatten_multimap(g_elt)),
            _calculate_linking_metric_core([(k, v) for k, v in p_lgh.items()], [(k, v) for k, v in g_lgh.items()]),
            _calculate_linking_metric_core([(k, v) for k, v in p_lgt.items()], [(k, v) for k, v in g_lgt.items()]),
        ]
        entry = {"fname": fname}
        row = [fname]
        for name, (p, r, f1, n_pred, n_gt, n_correct) in zip(_TASKS, res):
            entry[name] = {"num_pred": n_pred, "num_gt": n_gt, "num_correct": n_correct, "precision": p, "recall": r,
                           "f1": f1}
            row += [n_pred, n_gt, n_correct]
        entry["detail"] = kv[6]
        sample_details.append(entry)
        rows.append(row)
    totals, _processed = _accumulate(_gather_rows(rows, 18), 18)
    detail, flat = {}, {}
    for q, name in enumerate(_TASKS):
        n_pred, n_gt, n_correct = totals[3 * q : 3 * q + 3]
        p, r, f1 = _prf(n_correct, n_pred, n_gt)
        detail[name] = {"precision": p, "recall": r, "f1": f1, "num_pred": n_pred, "num_gt": n_gt, "num_correct": n_correct}
        prefix = "" if name == "kv_pair" else name + "_"
        flat[prefix + "precision"], flat[prefix + "recall"], flat[prefix + "f1"] = p, r, f1
    detail["detail"] = sample_details
    return flat, detail

"""Streaming evaluation loop — "next" row N3 of the scope table.

The reference's ``PEneoTrainer.prediction_loop`` (pipeline/trainer.py:57-211) keeps the five ``[P, C]`` logits tensors
of EVERY evaluation sample alive on the GPU until the loop ends (trainer.py:133-141: 7.3 MB per seq-512 document,
117 MB at seq 2048), then ``compute_metrics`` (start/run_rfund.py:243-304) decodes predictions and ground truth of the
whole set in one go.  Here each batch is decoded as soon as the model has produced it:

* ``StreamingEvaluator.add_batch`` enqueues the prediction decode (K3 + K4 on the batch's logits) and the ground-truth
  decode (same kernels on the tag tensors) without blocking, after which the logits can be freed;
* the host glue (strings, boxes, ordered dicts) of batch ``i`` runs while the GPU works on batch ``i + 1``;
* ``finish()`` returns exactly what ``decode_peneo`` returns for the accumulated set — ``(all_pred, all_gt, file_ids)``
  — so ``calculate_KVPE_metric`` / ``calculate_detail_KVPE_metric`` and the ``detail.json`` dump are unchanged.

``prediction_loop`` is the drop-in for the body of ``PEneoTrainer.prediction_loop``: same inputs consumed from the
dataloader's batches (``text``, ``fname``, the five ``*_shaking_tag`` tensors), same metric keys, including the
reference's quirks (losses of the LAST batch only, trainer.py:185-200; ``line_grouping_h2h_loss`` overwritten by the
t2t value, trainer.py:198-200).
"""
from __future__ import annotations

from collections import deque
from typing import Callable, Dict, Iterable, List, Optional

import torch

from . import decode, evaluation

_OUT_KEYS = ("line_extraction_shaking_outputs", "ent_linking_h2h_shaking_outputs", "ent_linking_t2t_shaking_outputs",
             "line_grouping_h2h_shaking_outputs", "line_grouping_t2t_shaking_outputs")
_TAG_KEYS = ("line_extraction_shaking_tag", "ent_linking_head_rel_shaking_tag", "ent_linking_tail_rel_shaking_tag",
             "line_grouping_head_rel_shaking_tag", "line_grouping_tail_rel_shaking_tag")


class StreamingEvaluator:
    """Per-batch decode of predictions and ground truth with bounded GPU memory (at most ``depth`` batches of compact
    records in flight; no logits are retained)."""

    def __init__(self, depth: int = 2, score_thresh: float = 0.0, device=None):
        self.depth, self.score_thresh = depth, score_thresh
        self.device = torch.device(device) if device is not None else None
        self.d2h = None
        self._pending = deque()
        self.all_pred: List = []
        self.all_gt: List = []
        self.file_ids: List = []

    def add_batch(self, outputs, inputs: Dict) -> None:
        """``outputs``: the model's ``PEneoOutput`` (or any object / dict with the five ``*_shaking_outputs`` and
        ``orig_bbox``); ``inputs``: the collated batch (needs ``text`` and the five tag tensors; ``fname`` optional)."""
        get = (lambda k: outputs[k]) if isinstance(outputs, dict) else (lambda k: getattr(outputs, k))
        if "text" not in inputs:
            raise ValueError("No text given in evaluation")  # pipeline/trainer.py:125-129
        logits = [get(k) for k in _OUT_KEYS]
        dev = self.device or next(t.device for t in logits if t.is_cuda)
        if self.d2h is None:
            self.d2h = torch.cuda.Stream(dev)
        orig_bbox = get("orig_bbox")
        n = int(orig_bbox.shape[1]) if torch.is_tensor(orig_bbox) else len(orig_bbox[0])  # seq_len = len(orig_bbox[s])
        tags = [inputs[k] for k in _TAG_KEYS]
        tags = [t if t.is_cuda else t.to(dev, non_blocking=True) for t in tags]
        with torch.no_grad():
            pred = decode.device_decode_async([l.detach() for l in logits], n, score_thresh=self.score_thresh, d2h_stream=self.d2h)
            gt = decode.device_decode_async(tags, n, decode_gt=True, d2h_stream=self.d2h)
        self._pending.append((pred, gt, list(inputs["text"]), list(inputs.get("fname", []))))
        while len(self._pending) > self.depth:
            self._drain_one()

    def _drain_one(self) -> None:
        pred, gt, texts, fnames = self._pending.popleft()
        dp, dg = pred.finish(), gt.finish()
        rows = range(dp.batch)
        self.all_pred += decode.assemble_many(dp, rows, texts)
        self.all_gt += decode.assemble_many(dg, rows, texts)
        self.file_ids += fnames

    def finish(self):
        """-> ``(all_pred_results, all_gt_results, all_fname)`` as ``decode_peneo`` returns them."""
        while self._pending:
            self._drain_one()
        return self.all_pred, self.all_gt, self.file_ids

    def metrics(self, detail_eval: bool = False):
        """-> ``(metric, detail)`` of ``calculate_[detail_]KVPE_metric`` (gathered over ranks when distributed)."""
        pred, gt, names = self.finish()
        fn = evaluation.calculate_detail_KVPE_metric if detail_eval else evaluation.calculate_KVPE_metric
        return fn(all_pred=pred, all_gt=gt, all_fname=names)


def prediction_loop(model: Callable, dataloader: Iterable[Dict], metric_key_prefix: str = "eval", detail_eval: bool = False,
                    prepare_inputs: Optional[Callable[[Dict], Dict]] = None, depth: int = 2, return_detail: bool = False):
    """Evaluation loop with the reference's metric dictionary (pipeline/trainer.py:102-211 with the ``compute_metrics``
    of start/run_rfund.py:243-304 inlined), decoding batch by batch.  ``model(**inputs)`` must return a ``PEneoOutput``."""
    ev = StreamingEvaluator(depth=depth)
    outputs = None
    for inputs in dataloader:
        if prepare_inputs is not None:
            inputs = prepare_inputs(inputs)
        with torch.no_grad():
            outputs = model(**inputs)
        ev.add_batch(outputs, inputs)
    if outputs is None:
        raise ValueError("empty evaluation dataloader")
    metric, detail = ev.metrics(detail_eval)
    _metrics = dict(metric)
    _metrics[f"{metric_key_prefix}_loss"] = outputs.loss.mean().item()
    _metrics[f"{metric_key_prefix}_line_extraction_loss"] = outputs.line_extraction_loss.mean().item()
    _metrics[f"{metric_key_prefix}_ent_linking_h2h_loss"] = outputs.ent_linking_h2h_loss.mean().item()
    _metrics[f"{metric_key_prefix}_ent_linking_t2t_loss"] = outputs.ent_linking_t2t_loss.mean().item()
    _metrics[f"{metric_key_prefix}_line_grouping_h2h_loss"] = outputs.line_grouping_h2h_loss.mean().item()
    _metrics[f"{metric_key_prefix}_line_grouping_h2h_loss"] = outputs.line_grouping_t2t_loss.mean().item()  # sic (trainer.py:198-200)
    metrics = {}
    for key in list(_metrics.keys()):
        metrics[key if key.startswith(f"{metric_key_prefix}_") else f"{metric_key_prefix}_{key}"] = _metrics.pop(key)
    return (metrics, detail) if return_detail else metrics

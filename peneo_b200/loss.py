"""``CrossEntropyLossOHEM`` — drop-in for model/custom_loss.py:104-288 of ZeningLin/PEneo on the CUDA loss kernels.

Covers what ``PEneoDecoder`` can configure (model/peneo_decoder.py:304-313): class weights (2 or 3 classes),
``num_hard_positive`` / ``num_hard_negative`` (or the ``hard_*_ratio`` forms, including the constructor's
``hard_positive_ratio = hard_negative_ratio`` assignment, custom_loss.py:164), reduction ``"mean"``.  The random
pre-sampling mode (``random=True``, Python ``random.sample``) and the ``"sum"`` / ``"none"`` reductions are never
enabled by the decoder and are not implemented.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .autograd import _PairLoss
from .ohem import _OhemLoss
from .ops import HEAD_CLASSES


class CrossEntropyLossOHEM(nn.Module):
    def __init__(self, hard_positive_ratio: float = None, hard_negative_ratio: float = None, num_hard_positive: int = -1,
                 num_hard_negative: int = -1, weight: Optional[torch.Tensor] = None, size_average=None,
                 ignore_index: int = -100, reduction: str = "mean", random: bool = False) -> None:
        super().__init__()
        if weight is None or weight.numel() not in (2, 3):
            raise NotImplementedError("the CUDA loss needs class weights for 2 (line extraction) or 3 (link) classes")
        self.register_buffer("weight", weight.detach().clone().float())
        self.ignore_index, self.reduction = ignore_index, reduction
        hard_positive_ratio = hard_positive_ratio or None
        hard_negative_ratio = hard_negative_ratio or None
        if hard_positive_ratio is not None:
            assert 0 <= int(hard_positive_ratio * 1000) <= 1000, f"hard_positive_ratio must be in [0, 1], {hard_positive_ratio} given"
            self.num_hard_positive, self.hard_positive_ratio = None, hard_negative_ratio  # sic (custom_loss.py:164)
        elif num_hard_positive is not None:
            self.num_hard_positive, self.hard_positive_ratio = num_hard_positive, None
        else:
            raise ValueError("either num_hard_positive or hard_positive_ratio must be given")
        if hard_negative_ratio is not None:
            assert 0 <= int(hard_negative_ratio * 1000) <= 1000, f"hard_negative_ratio must be in [0, 1], {hard_negative_ratio} given"
            self.num_hard_negative, self.hard_negative_ratio = None, hard_negative_ratio
        elif num_hard_negative is not None:
            self.num_hard_negative, self.hard_negative_ratio = num_hard_negative, None
        else:
            raise ValueError("either num_hard_negative or hard_negative_ratio must be given")
        if random:
            raise NotImplementedError("random pre-sampling (random=True) is not implemented on the CUDA path")
        self.random = False

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if self.reduction not in ("none", "mean", "sum"):
            raise ValueError(f"the given reduction value {self.reduction} is invalid, must be 'none', 'mean' or 'sum' ")
        if self.reduction != "mean":
            raise NotImplementedError("only reduction='mean' runs on the CUDA path")
        c = input.shape[-1]
        if c != self.weight.numel():
            raise ValueError(f"{c} classes but {self.weight.numel()} class weights")
        x = input.reshape(-1, 1, c)  # M elements = M "documents" of one pair each
        t = target.reshape(-1, 1)
        m = x.shape[0]
        head = 0 if c == 2 else 1
        logits = [x if h == head else torch.zeros(m, 1, ch, dtype=torch.float32, device=x.device)
                  for h, ch in enumerate(HEAD_CLASSES)]
        zeros = torch.zeros(m, 1, dtype=torch.int64, device=x.device)
        tags = [t if h == head else zeros for h in range(5)]
        w = self.weight.tolist()
        ohem_off = (self.num_hard_positive == -1 and self.num_hard_negative == -1 and self.hard_positive_ratio is None
                    and self.hard_negative_ratio is None)
        if ohem_off:
            return _PairLoss.apply(w + [0.0] * (3 - len(w)), None, *tags, *logits)[head]
        if self.hard_positive_ratio is not None:  # custom_loss.py:212-222
            hp = int(m * self.hard_positive_ratio)
            self.num_hard_positive = hp if hp != 0 else m
        if self.hard_negative_ratio is not None:  # custom_loss.py:224-233
            hn = int(m * self.hard_negative_ratio)
            self.num_hard_negative = hn if hn != 0 else m
        return _OhemLoss.apply(w, (self.num_hard_positive, self.num_hard_negative), *tags, *logits)[head]

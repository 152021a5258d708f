"""autograd glue: the CUDA kernels behind ``torch.autograd.Function`` so that the decoder trains
inside the reference's Trainer / DDP exactly like the module it replaces."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops


class _PairLoss(torch.autograd.Function):
    """Five class-weighted CE sub-losses + their ratio-weighted sum
    (model/peneo_decoder.py:375-428, model/custom_loss.py:189-202), one pass over the logits."""

    @staticmethod
    def forward(ctx, class_w, ratios, t0, t1, t2, t3, t4, l0, l1, l2, l3, l4):
        out6, saved = ops.pair_loss([l0, l1, l2, l3, l4], [t0, t1, t2, t3, t4], class_w, ratios)
        ctx.saved = saved
        return out6

    @staticmethod
    def backward(ctx, g6):
        # d(sum_h g6[h] * loss_h + g6[5] * sum_h ratio_h loss_h) / d logits_h: the per-head scale is applied inside
        # the kernel (no elementwise passes over the [B, P, C] gradients on the torch side)
        dl = ops.pair_loss_backward(ctx.saved, g6)
        return (None, None, None, None, None, None, None, *dl)


def pair_loss_op(logits: Sequence[torch.Tensor], tags: Sequence[torch.Tensor], class_w: Sequence[float],
                 ratios: Optional[Sequence[float]]):
    out6 = _PairLoss.apply(list(class_w), None if ratios is None else list(ratios), *tags, *logits)
    return out6[5], [out6[h] for h in range(5)]


def decoder_forward_with_grad(decoder, sequence_output: torch.Tensor, dropout=None) -> List[torch.Tensor]:
    from .train import heads_with_grad

    return heads_with_grad(decoder, sequence_output, dropout)

"""OHEM variant of the pairwise loss (model/custom_loss.py:204-288) behind autograd.

``PEneoDecoder`` passes only ``num_hard_positive`` / ``num_hard_negative`` to its two
``CrossEntropyLossOHEM`` objects (model/peneo_decoder.py:304-313); selection, its quirks and the
divisor are reproduced in ``csrc/ohem.cu``.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from . import _lib
from .ops import COUNTERS, NUM_HEADS, _require_cuda, _stream, seq_len_from_pairs


class _OhemLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, class_w, ohem, t0, t1, t2, t3, t4, l0, l1, l2, l3, l4):
        lib = _lib.load()
        logits, tags = [l0, l1, l2, l3, l4], [t0, t1, t2, t3, t4]
        b, p, _ = logits[0].shape
        n = seq_len_from_pairs(p)
        dev = logits[0].device
        lg = [l.detach() if (l.dtype == torch.float32 and l.is_contiguous()) else l.detach().float().contiguous()
              for l in logits]
        tg = []
        for t in tags:
            _require_cuda(t, "shaking tag")
            if tuple(t.shape) != (b, p):
                raise AssertionError("invalid input shape")  # model/peneo_decoder.py:329-331
            tg.append(t if (t.dtype == torch.int64 and t.is_contiguous()) else t.long().contiguous())
        w3 = list(class_w) + [0.0] * (3 - len(class_w))
        out6 = torch.empty(6, dtype=torch.float32, device=dev)
        ws = torch.empty(lib.peneo_pair_loss_ohem_workspace_bytes(b, n), dtype=torch.uint8, device=dev)
        COUNTERS["kernels"] += 1
        _lib.check(
            lib.peneo_pair_loss_ohem_fwd(b, n, _lib.ptrs5(lg), _lib.ptrs5(tg), _lib.floats(w3), _lib.floats([1.0] * 5),
                                         int(ohem[0]), int(ohem[1]), out6.data_ptr(), ws.data_ptr(), _stream(dev)),
            "peneo_pair_loss_ohem_fwd",
        )
        ctx.saved = (ws, lg, tg, w3, b, n)
        return out6[:5].clone()

    @staticmethod
    def backward(ctx, g5):
        lib = _lib.load()
        ws, lg, tg, w3, b, n = ctx.saved
        dev = lg[0].device
        g6 = torch.zeros(6, dtype=torch.float32, device=dev)  # d L / d (five sub-losses, unused total)
        g6[:5] = g5.detach().float()
        dl = [torch.empty_like(l) for l in lg]
        _lib.check(
            lib.peneo_pair_loss_ohem_bwd(b, n, _lib.ptrs5(lg), _lib.ptrs5(tg), _lib.floats(w3), _lib.floats([1.0] * 5),
                                         g6.data_ptr(), ws.data_ptr(), _lib.ptrs5(dl), _stream(dev)),
            "peneo_pair_loss_ohem_bwd",
        )
        return (None, None, None, None, None, None, None, *dl)


def ohem_losses(logits: Sequence[torch.Tensor], tags: Sequence[torch.Tensor], class_w: torch.Tensor,
                ohem: Tuple[int, int]) -> List[torch.Tensor]:
    """Five OHEM sub-losses (LE uses the first two class weights, model/peneo_decoder.py:302-303)."""
    subs = _OhemLoss.apply([float(v) for v in class_w.tolist()], tuple(ohem), *tags, *logits)
    return [subs[h] for h in range(NUM_HEADS)]

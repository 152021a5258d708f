"""``HandshakingTaggingScheme`` — drop-in for model/peneo_decoder.py:12-115 backed by the CUDA
scatter / spot-extraction kernels."""
from __future__ import annotations

from typing import List, Tuple

import torch

from . import ops
from .decode import device_decode, spots_from_device


class HandshakingTaggingScheme:
    @staticmethod
    def spots2shaking_tag4batch(batch_spots, shaking_ind2matrix_ind=None, matrix_ind2shaking_ind=None,
                                seq_len: int = None, device=None) -> torch.Tensor:
        """Dense [B, P] int64 tags from per-sample spot lists [(i, j, tag), ...]
        (model/peneo_decoder.py:34-73).  Built on the GPU; returned on ``device`` (default: CPU,
        where the reference's collator expects it — pass ``device="cuda"`` to skip the copy)."""
        if shaking_ind2matrix_ind is not None and matrix_ind2shaking_ind is not None:
            seq_len = len(matrix_ind2shaking_ind)
        elif seq_len is None:
            raise ValueError(
                "If shaking_ind2matrix_ind and matrix_ind2shaking_ind are not provided,seq_len must be given"
            )
        quads = [(b, sp[0], sp[1], sp[2]) for b, spots in enumerate(batch_spots) for sp in spots]
        q = torch.tensor(quads, dtype=torch.int64).reshape(-1, 4)
        if q.numel():
            # the reference indexes an N x N Python table with the spot: indices in [-N, N) are legal (negative ones
            # count from the end), anything else raises IndexError (model/peneo_decoder.py:70).  Same here, on the
            # host, before the kernel sees the list; spots below the diagonal go to cell 0 inside the kernel.
            ij = q[:, 1:3]
            if bool(((ij < -seq_len) | (ij >= seq_len)).any()):
                raise IndexError("list index out of range")
            q[:, 1:3] = torch.where(ij < 0, ij + seq_len, ij)
        q = q.to(torch.int32).cuda()
        tags = ops.scatter_tags(q, len(batch_spots), seq_len)
        return tags.cpu() if device is None else tags.to(device)

    @staticmethod
    def get_spots_from_shaking_tag(shaking_tag: torch.Tensor, shaking_ind2matrix_ind=None,
                                   seq_len: int = None) -> List[Tuple]:
        """[(i, j, tag, score)] for every non-zero prediction in increasing flat index
        (model/peneo_decoder.py:75-115)."""
        if shaking_ind2matrix_ind is not None:
            seq_len = ops.seq_len_from_pairs(len(shaking_ind2matrix_ind))
        elif seq_len is None:
            raise ValueError(
                "If shaking_ind2matrix_ind and matrix_ind2shaking_ind are not provided,seq_len must be given"
            )
        t = shaking_tag if shaking_tag.is_cuda else shaking_tag.cuda()
        tag_mode = not (t.dim() > 1 and t.shape[-1] > 1)
        if tag_mode:
            five, head = [t.reshape(1, -1)] * 5, 0
        elif t.shape[-1] == 2:
            dummy = torch.zeros(1, t.shape[0], 3, dtype=t.dtype, device=t.device)
            five, head = [t.unsqueeze(0), dummy, dummy, dummy, dummy], 0
        else:
            dummy = torch.zeros(1, t.shape[0], 2, dtype=t.dtype, device=t.device)
            five, head = [dummy] + [t.unsqueeze(0)] * 4, 1
        dd = device_decode(five, seq_len, decode_gt=tag_mode, want_spots=True)
        return spots_from_device(dd, 0, head)

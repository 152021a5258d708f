"""Functional wrappers over the C ABI: torch tensors in, torch tensors out, everything enqueued
on the current CUDA stream of the input's device.  torch is used for device memory and streams
only; all arithmetic happens in libpeneo_b200.so.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from ._lib import DT_BF16, DT_F16, DT_F32, DT_I64, HEAD_CLASSES, NUM_HEADS, PREC_BF16, PREC_FP32

HEAD_NAMES = ("line_extraction", "ent_linking_h2h", "ent_linking_t2t", "line_grouping_h2h", "line_grouping_t2t")

# our own kernel launches since import (bench.py reports the per-step delta as gpu_launches)
COUNTERS = {"kernels": 0}

_TORCH_DT = {torch.float32: DT_F32, torch.bfloat16: DT_BF16, torch.float16: DT_F16, torch.int64: DT_I64}


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor: peneo_b200 has no CPU path")


def shaking_len(n: int) -> int:
    return n * (n + 1) // 2


def seq_len_from_pairs(p: int) -> int:
    n = int(((8 * p + 1) ** 0.5 - 1) / 2)
    while shaking_len(n) < p:
        n += 1
    while shaking_len(n) > p:
        n -= 1
    if shaking_len(n) != p:
        raise ValueError(f"{p} is not a triangular number")
    return n


@dataclass
class DecoderDims:
    hin: int
    hid: int
    d: int
    shrink: bool
    num_layers: int

    def c(self) -> _lib.Dims:
        return _lib.Dims(self.hin, self.hid, self.d, int(self.shrink), self.num_layers)

    def bf16_capable(self) -> bool:
        """The fused tcgen05 path (K2 / T1, forward and backward): the shipped configuration."""
        return self.shrink and self.hid == 768 and self.d == 384 and self.num_layers == 2 and self.hin % 64 == 0

    def bf16_forward_capable(self) -> bool:
        """PREC_BF16 is accepted for the forward pass: the fused path, or the unfused tensor-core forward
        (csrc/pair_heads_generic.cu) for any widths in multiples of 64 and up to 8 classifier layers."""
        if self.bf16_capable():
            return True
        return (self.d % 64 == 0 and self.hin % 64 == 0 and (not self.shrink or self.hid % 64 == 0)
                and 1 <= self.num_layers <= 8)


class WeightPack:
    """Device buffer with the kernel-layout weights for one (dims, precision)."""

    def __init__(self, dims: DecoderDims, prec: int, device):
        lib = _lib.load()
        self.dims, self.prec = dims, prec
        nbytes = lib.peneo_pack_bytes(dims.c(), prec)
        if nbytes == 0:
            raise RuntimeError(f"unsupported configuration for precision {prec}: {lib.peneo_last_error().decode()}")
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)

    def update(self, state: Dict[str, torch.Tensor]) -> None:
        """(Re)pack from fp32 CUDA tensors keyed like the reference state dict."""
        lib = _lib.load()
        dm = self.dims
        keep = []  # keep converted tensors alive until the pack kernels are enqueued

        def ptr(key):
            t = state[key]
            _require_cuda(t, key)
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.detach().to(torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()

        P = _lib.Params()
        if dm.shrink:
            P.shrink_w1, P.shrink_b1 = ptr("shrink_projection.0.weight"), ptr("shrink_projection.0.bias")
            P.shrink_w2, P.shrink_b2 = ptr("shrink_projection.3.weight"), ptr("shrink_projection.3.bias")
        P.combine_w, P.combine_b = ptr("handshaking_kernel.combine_fc.weight"), ptr("handshaking_kernel.combine_fc.bias")
        for h, name in enumerate(HEAD_NAMES):
            if dm.num_layers == 1:
                P.out_w[h], P.out_b[h] = ptr(f"{name}_fc.weight"), ptr(f"{name}_fc.bias")
            else:
                for l in range(dm.num_layers - 1):
                    P.mid_w[h * 8 + l], P.mid_b[h * 8 + l] = ptr(f"{name}_fc.{3 * l}.weight"), ptr(f"{name}_fc.{3 * l}.bias")
                last = 3 * (dm.num_layers - 1)
                P.out_w[h], P.out_b[h] = ptr(f"{name}_fc.{last}.weight"), ptr(f"{name}_fc.{last}.bias")
        _lib.check(lib.peneo_pack_weights(dm.c(), self.prec, P, self.buf.data_ptr(), _stream(self.buf.device)),
                   "peneo_pack_weights")
        del keep


def seam_tokens(x: torch.Tensor, out_dtype: torch.dtype, in_dropout=None) -> torch.Tensor:
    """The backbone -> decoder seam (model/modeling_peneo.py:134-165) as at most ONE pass: ``x`` is ``[B, N, hin]``, usually
    a strided view of the backbone output with the CLS row / visual tokens stripped.  Returns a ``[B * N, hin]`` matrix
    with a uniform row stride that the kernels read in place: a *view* when ``x`` already is one (B == 1 or a dense batch)
    in ``out_dtype`` and no input dropout is asked for, otherwise the output of ``peneo_gather_tokens`` (strip + dropout
    + cast in one kernel — no intermediate ``.contiguous()`` / dropout / cast tensors).  ``in_dropout``: None or (p, seed)."""
    lib = _lib.load()
    _require_cuda(x, "sequence_output")
    if x.dtype not in (torch.float32, torch.bfloat16, torch.float16):
        x = x.float()
    if x.dim() == 2:
        x = x.unsqueeze(0)
    b, n, hin = x.shape
    if x.stride(2) != 1:
        x = x.contiguous()
    drop_on = in_dropout is not None and in_dropout[0] > 0.0
    uniform = b == 1 or x.stride(0) == n * x.stride(1)
    if uniform and not drop_on and x.dtype == out_dtype:
        return x.as_strided((b * n, hin), (x.stride(1), 1))
    out = torch.empty(b * n, hin, dtype=out_dtype, device=x.device)
    COUNTERS["kernels"] += 1
    _lib.check(
        lib.peneo_gather_tokens(x.data_ptr(), _TORCH_DT[x.dtype], b, n, hin, x.stride(0), x.stride(1), out.data_ptr(),
                                _TORCH_DT[out_dtype], _lib.dropout_arg(in_dropout), _stream(x.device)),
        "peneo_gather_tokens",
    )
    return out


def token_projections(pack: WeightPack, x: torch.Tensor, dropout=None, in_dropout=None) -> torch.Tensor:
    """x: [..., hin] (fp32 / bf16 / fp16, last dim contiguous; a strided [B, N, hin] view is consumed in place or through
    one gather pass, see :func:`seam_tokens`) -> ab [tokens, 2d] (fp32, or bf16 pre-multiplied by 1/2 in bf16 mode).
    ``dropout``: the decoder's own Dropout modules, None (eval) or ``(p, seed)``; ``in_dropout``: dropout on x itself."""
    lib = _lib.load()
    _require_cuda(x, "sequence_output")
    dm = pack.dims
    if x.shape[-1] != dm.hin:
        raise ValueError(f"last dimension {x.shape[-1]} != input_size {dm.hin}")
    if x.dim() == 3:
        x2 = seam_tokens(x, torch.float32 if pack.prec == PREC_FP32 else torch.bfloat16, in_dropout)
    else:
        if x.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            x = x.float()
        x2 = x.reshape(-1, dm.hin)
        if x2.stride(-1) != 1:
            x2 = x2.contiguous()
    tokens = x2.shape[0]
    out_dt = torch.float32 if pack.prec == PREC_FP32 else torch.bfloat16
    ab = torch.empty(tokens, 2 * dm.d, dtype=out_dt, device=x.device)
    direct = (x2.dtype == out_dt and x2.stride(0) % (8 if pack.prec == PREC_BF16 else 4) == 0 and x2.data_ptr() % 16 == 0)
    COUNTERS["kernels"] += (0 if direct else 1) + (4 if pack.prec == PREC_FP32 and dm.shrink else 3 if dm.shrink else 2)
    ws = torch.empty(lib.peneo_token_proj_workspace_bytes(dm.c(), pack.prec, tokens), dtype=torch.uint8, device=x.device)
    _lib.check(
        lib.peneo_token_proj_fwd(dm.c(), pack.prec, pack.buf.data_ptr(), x2.data_ptr(), _TORCH_DT[x2.dtype],
                                 x2.stride(0) if tokens > 1 else dm.hin, tokens, ab.data_ptr(), ws.data_ptr(),
                                 _lib.dropout_arg(dropout), _stream(x.device)),
        "peneo_token_proj_fwd",
    )
    return ab


def pair_heads(pack: WeightPack, ab: torch.Tensor, batch: int, n: int, dropout=None) -> List[torch.Tensor]:
    """ab from :func:`token_projections` -> five fp32 logits tensors [batch, P, C_h]."""
    lib = _lib.load()
    p = shaking_len(n)
    logits = [torch.empty(batch, p, c, dtype=torch.float32, device=ab.device) for c in HEAD_CLASSES]
    COUNTERS["kernels"] += 1
    _lib.check(
        lib.peneo_pair_heads_fwd(pack.dims.c(), pack.prec, pack.buf.data_ptr(), ab.data_ptr(), batch, n,
                                 _lib.ptrs5(logits), _lib.dropout_arg(dropout), _stream(ab.device)),
        "peneo_pair_heads_fwd",
    )
    return logits


def heads_forward(pack: WeightPack, x: torch.Tensor, dropout=None, in_dropout=None) -> List[torch.Tensor]:
    """[B, N, hin] hidden states -> five logits tensors (return order LE, ELh, ELt, LGh, LGt)."""
    if x.dim() != 3:
        raise ValueError("sequence_output must be [batch, seq_len, hidden]")
    b, n, _ = x.shape
    return pair_heads(pack, token_projections(pack, x, dropout, in_dropout), b, n, dropout)


def _check_tags(tags: Sequence[torch.Tensor], b: int, p: int) -> List[torch.Tensor]:
    tg = []
    for k in range(NUM_HEADS):
        t = tags[k]
        _require_cuda(t, "shaking tag")
        if tuple(t.shape) != (b, p):
            raise AssertionError("invalid input shape")  # model/peneo_decoder.py:329-331
        tg.append(t if (t.dtype == torch.int64 and t.is_contiguous()) else t.long().contiguous())
    return tg


def saved_activation_budget_bytes() -> int:
    """How much the training forward may keep for the backward (h + s, 4.5 KB per pair).  ``PENEO_SAVE_ACT_GB`` (0 =
    never save: the backward regenerates the activations on the tensor cores); default 24 GB."""
    import os

    return int(float(os.environ.get("PENEO_SAVE_ACT_GB", "24")) * 2**30)


def heads_loss_forward(pack: WeightPack, x: torch.Tensor, tags: Sequence[torch.Tensor], class_weights: Sequence[float],
                       ratios: Optional[Sequence[float]] = None, dropout=None, in_dropout=None, save: bool = False):
    """Heads + class-weighted CE in one sweep over the pair tiles (``peneo_pair_heads_loss_fwd``): the K2 epilogue that
    writes a pair's logits also reduces its loss terms.  Returns (logits, out6, ctx); ctx feeds the fused backward.
    ``save``: also keep the hidden pre-activations and the pair representations (bf16, 4.5 KB per pair) when they fit
    the budget, so that the backward pass needs no recompute GEMM."""
    lib = _lib.load()
    if x.dim() != 3:
        raise ValueError("sequence_output must be [batch, seq_len, hidden]")
    b, n, _ = x.shape
    p = shaking_len(n)
    tg = _check_tags(tags, b, p)
    ab = token_projections(pack, x, dropout, in_dropout)
    logits = [torch.empty(b, p, c, dtype=torch.float32, device=ab.device) for c in HEAD_CLASSES]
    out6 = torch.empty(6, dtype=torch.float32, device=ab.device)
    ws = torch.empty(lib.peneo_pair_loss_workspace_bytes(b, n), dtype=torch.uint8, device=ab.device)
    w3 = list(class_weights) + [0.0] * (3 - len(class_weights))
    r5 = [1.0] * 5 if ratios is None else list(ratios)
    COUNTERS["kernels"] += 2
    saved = None
    d = pack.dims.d
    if save and b * p * 6 * d * 2 <= saved_activation_budget_bytes():
        try:
            saved = (torch.empty(b * p, 5 * d, dtype=torch.bfloat16, device=ab.device),
                     torch.empty(b * p, d, dtype=torch.bfloat16, device=ab.device))
        except torch.cuda.OutOfMemoryError:  # no room next to the backbone's activations: regenerate in the backward pass
            saved = None
    sa = _lib.SavedAct(saved[0].data_ptr(), saved[1].data_ptr()) if saved is not None else None
    _lib.check(
        lib.peneo_pair_heads_loss_fwd(pack.dims.c(), pack.prec, pack.buf.data_ptr(), ab.data_ptr(), b, n, _lib.ptrs5(logits),
                                      _lib.ptrs5(tg), _lib.floats(w3), _lib.floats(r5), out6.data_ptr(), ws.data_ptr(),
                                      _lib.dropout_arg(dropout), _stream(ab.device), sa),
        "peneo_pair_heads_loss_fwd",
    )
    return logits, out6, (ws, tg, w3, r5, saved)


def pair_loss(logits: Sequence[torch.Tensor], tags: Sequence[torch.Tensor], class_weights: Sequence[float],
              ratios: Optional[Sequence[float]] = None):
    """Weighted-mean CE of the five heads (OHEM off).  Returns (out6, workspace): out6[0:5] are the
    sub-losses, out6[5] their ratio-weighted sum; workspace feeds :func:`pair_loss_backward`."""
    lib = _lib.load()
    b, p, _ = logits[0].shape
    n = seq_len_from_pairs(p)
    dev = logits[0].device
    lg = [l if (l.dtype == torch.float32 and l.is_contiguous()) else l.float().contiguous() for l in logits]
    tg = []
    for k in range(NUM_HEADS):
        t = tags[k]
        _require_cuda(t, "shaking tag")
        if tuple(t.shape) != (b, p):
            raise AssertionError("invalid input shape")  # model/peneo_decoder.py:329-331
        tg.append(t if (t.dtype == torch.int64 and t.is_contiguous()) else t.long().contiguous())
    out6 = torch.empty(6, dtype=torch.float32, device=dev)
    ws = torch.empty(lib.peneo_pair_loss_workspace_bytes(b, n), dtype=torch.uint8, device=dev)
    w3 = list(class_weights) + [0.0] * (3 - len(class_weights))
    r5 = [1.0] * 5 if ratios is None else list(ratios)
    _lib.check(
        lib.peneo_pair_loss_fwd(b, n, _lib.ptrs5(lg), _lib.ptrs5(tg), _lib.floats(w3), _lib.floats(r5),
                                out6.data_ptr(), ws.data_ptr(), _stream(dev)),
        "peneo_pair_loss_fwd",
    )
    return out6, (ws, lg, tg, w3, r5, b, n)


def pair_loss_backward(ctx, grad_out6: torch.Tensor) -> List[torch.Tensor]:
    """grad_out6: d L / d out6 of :func:`pair_loss` (six values) -> d L / d logits of the five heads."""
    lib = _lib.load()
    ws, lg, tg, w3, r5, b, n = ctx
    dev = lg[0].device
    g = grad_out6.detach().to(device=dev, dtype=torch.float32).reshape(6).contiguous()
    dl = [torch.empty_like(l) for l in lg]
    _lib.check(
        lib.peneo_pair_loss_bwd(b, n, _lib.ptrs5(lg), _lib.ptrs5(tg), _lib.floats(w3), _lib.floats(r5), g.data_ptr(),
                                ws.data_ptr(), _lib.ptrs5(dl), _stream(dev)),
        "peneo_pair_loss_bwd",
    )
    return dl


def scatter_tags(spots_bijt: torch.Tensor, batch: int, n: int) -> torch.Tensor:
    """int32 [S, 4] (doc, i, j, tag) quadruples on the GPU -> dense int64 [batch, P] tags."""
    lib = _lib.load()
    _require_cuda(spots_bijt, "spots")
    sp = spots_bijt.to(torch.int32).contiguous()
    tags = torch.empty(batch, shaking_len(n), dtype=torch.int64, device=sp.device)
    _lib.check(lib.peneo_scatter_tags(sp.data_ptr(), sp.shape[0], batch, n, tags.data_ptr(), _stream(sp.device)),
               "peneo_scatter_tags")
    return tags


def selftest() -> (int, str):
    import ctypes as C

    lib = _lib.load()
    mask = C.c_uint32(0)
    buf = C.create_string_buffer(4096)
    _lib.check(lib.peneo_selftest(C.byref(mask), buf, 4096), "peneo_selftest")
    return mask.value, buf.value.decode()


def probe_rates() -> Dict[str, float]:
    import ctypes as C

    lib = _lib.load()
    out = (C.c_double * 5)()
    _lib.check(lib.peneo_probe_rates(out, 5), "peneo_probe_rates")
    names = ("tanh_f32", "ex2_f32", "tanh_bf16x2_elems", "ffma_dependent", "silu_1mufu")
    return {k: out[i] for i, k in enumerate(names)}

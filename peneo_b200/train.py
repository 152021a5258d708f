"""Training path of ``PEneoDecoderB200``: one ``torch.autograd.Function`` around the whole decoder.

Forward runs the inference kernels (K1 + K2) and returns the five logits tensors; nothing of size
``[P, D]`` is saved.  Backward receives d loss / d logits (from the fused loss's own backward, plus
whatever else the caller hung on the logits), and calls ``peneo_heads_bwd`` which recomputes the pair
activations chunk by chunk and produces every parameter gradient and d sequence_output — what
autograd does for model/peneo_decoder.py:349-363 in the reference.
"""
from __future__ import annotations

from typing import Dict, List

import torch

from . import _lib, ops
from ._lib import PREC_BF16, PREC_FP32, PREC_TF32
from .ops import HEAD_NAMES, WeightPack


def _param_items(decoder) -> List:
    return [(k, p) for k, p in decoder.named_parameters()]


def _grad_struct(dims, grads: Dict[str, torch.Tensor]) -> _lib.Grads:
    g = _lib.Grads()
    ptr = lambda key: grads[key].data_ptr()  # noqa: E731
    if dims.shrink:
        g.shrink_w1, g.shrink_b1 = ptr("shrink_projection.0.weight"), ptr("shrink_projection.0.bias")
        g.shrink_w2, g.shrink_b2 = ptr("shrink_projection.3.weight"), ptr("shrink_projection.3.bias")
    g.combine_w, g.combine_b = ptr("handshaking_kernel.combine_fc.weight"), ptr("handshaking_kernel.combine_fc.bias")
    for h, name in enumerate(HEAD_NAMES):
        if dims.num_layers == 1:
            g.out_w[h], g.out_b[h] = ptr(f"{name}_fc.weight"), ptr(f"{name}_fc.bias")
        else:
            for l in range(dims.num_layers - 1):
                g.mid_w[h * 8 + l], g.mid_b[h * 8 + l] = ptr(f"{name}_fc.{3 * l}.weight"), ptr(f"{name}_fc.{3 * l}.bias")
            last = 3 * (dims.num_layers - 1)
            g.out_w[h], g.out_b[h] = ptr(f"{name}_fc.{last}.weight"), ptr(f"{name}_fc.{last}.bias")
    return g


def heads_backward(decoder, x: torch.Tensor, dlogits: List[torch.Tensor], need_dx: bool, dropout=None, fused=None,
                   in_dropout=None):
    """x: [B, N, hin] CUDA; dlogits: five fp32 [B, P, C_h]; dropout: the forward pass's (p, seed) or None.
    ``fused`` = (logits, loss ctx, grad_out6): the loss backward happens inside the pair tiles (``dlogits`` unused).
    Returns ({param_key: grad}, dx | None)."""
    lib = _lib.load()
    dims = decoder.dims
    b, n, hin = x.shape
    dev = x.device
    # bf16 mode: the [rows x 384] GEMMs of the pair part run on tcgen05 (bf16 operands, fp32 accumulation);
    # `backward_precision = "fp32"` on the module forces the exact CUDA-core path.
    # Configurations outside the fused kernels (shrink off, d != 384, num_layers != 2) keep fp32 buffers but run every
    # recompute / gradient GEMM on tcgen05 kind::tf32 (PREC_TF32, fp32 pack).
    if decoder.precision == "bf16" and getattr(decoder, "backward_precision", "bf16") == "bf16":
        if dims.bf16_capable():
            pack, prec = decoder._weight_pack(dev), PREC_BF16
        else:
            pack, prec = decoder._fp32_pack(dev), PREC_TF32
    else:
        pack, prec = decoder._fp32_pack(dev), PREC_FP32
    grads = {k: torch.empty_like(p, dtype=torch.float32) for k, p in _param_items(decoder)}
    # the same seam as the forward pass (strided view in place, or strip + input dropout + cast in one gather pass)
    xd = x.detach()
    x2 = ops.seam_tokens(xd, torch.float32 if xd.dtype not in (torch.bfloat16,) else torch.bfloat16, in_dropout)
    dx = torch.empty(b * n, hin, dtype=torch.float32, device=dev) if need_dx else None
    ws = torch.empty(lib.peneo_heads_bwd_workspace_bytes(dims.c(), prec, b, n), dtype=torch.uint8, device=dev)
    gs = _grad_struct(dims, grads)
    ops.COUNTERS["kernels"] += 1
    if fused is not None:
        logits, (lws, tg, w3, r5, saved), g6 = fused
        g6 = g6.detach().to(device=dev, dtype=torch.float32).reshape(6).contiguous()
        # saved = (h, s) of the forward pass, consumed here (h becomes G in place): a second backward through the same
        # graph finds them gone and regenerates the activations instead
        sa = _lib.SavedAct(saved[0].data_ptr(), saved[1].data_ptr()) if saved is not None else None
        _lib.check(
            lib.peneo_heads_loss_bwd(dims.c(), prec, pack.buf.data_ptr(), x2.data_ptr(), ops._TORCH_DT[x2.dtype],
                                     x2.stride(0) if b * n > 1 else hin, b, n, _lib.ptrs5(logits), _lib.ptrs5(tg),
                                     _lib.floats(w3), _lib.floats(r5), g6.data_ptr(), lws.data_ptr(), gs,
                                     dx.data_ptr() if dx is not None else None, ws.data_ptr(), _lib.dropout_arg(dropout),
                                     ops._stream(dev), sa),
            "peneo_heads_loss_bwd",
        )
        _mask_dx(lib, dx, b * n, hin, in_dropout, dev)
        return grads, (dx.view(b, n, hin) if dx is not None else None)
    dl = [g if (g.dtype == torch.float32 and g.is_contiguous()) else g.float().contiguous() for g in dlogits]
    _lib.check(
        lib.peneo_heads_bwd(dims.c(), prec, pack.buf.data_ptr(), x2.data_ptr(), ops._TORCH_DT[x2.dtype],
                            x2.stride(0) if b * n > 1 else hin, b, n, _lib.ptrs5(dl), gs,
                            dx.data_ptr() if dx is not None else None, ws.data_ptr(), _lib.dropout_arg(dropout),
                            ops._stream(dev)),
        "peneo_heads_bwd",
    )
    _mask_dx(lib, dx, b * n, hin, in_dropout, dev)
    return grads, (dx.view(b, n, hin) if dx is not None else None)


def _mask_dx(lib, dx, tokens, hin, in_dropout, dev):
    """backward of the input dropout fused into the seam: the regenerated mask on d loss / d sequence_output"""
    if dx is not None and in_dropout is not None and in_dropout[0] > 0.0:
        _lib.check(lib.peneo_token_dropout_bwd(dx.data_ptr(), tokens, hin, _lib.dropout_arg(in_dropout), ops._stream(dev)),
                   "peneo_token_dropout_bwd")


class _DecoderHeads(torch.autograd.Function):
    @staticmethod
    def forward(ctx, decoder, x, drop, *params):
        pack = decoder._weight_pack(x.device)
        ctx.dropout = drop  # (p, seed) of this step: the backward pass regenerates the same masks
        ctx.in_dropout = decoder._input_dropout(drop)
        logits = ops.heads_forward(pack, x.detach(), drop, ctx.in_dropout)
        ctx.decoder = decoder
        ctx.save_for_backward(x)
        ctx.need_dx = x.requires_grad
        ctx.param_keys = [k for k, _ in _param_items(decoder)]
        return tuple(logits)

    @staticmethod
    def backward(ctx, *glogits):
        (x,) = ctx.saved_tensors
        dec = ctx.decoder
        shapes = [(x.shape[0], ops.shaking_len(x.shape[1]), c) for c in ops.HEAD_CLASSES]
        dl = [g if g is not None else torch.zeros(s, dtype=torch.float32, device=x.device)
              for g, s in zip(glogits, shapes)]
        grads, dx = heads_backward(dec, x, dl, ctx.need_dx, ctx.dropout, in_dropout=ctx.in_dropout)
        out = [None, dx.to(x.dtype) if dx is not None else None, None]
        for k, p in _param_items(dec):
            out.append(grads[k].to(p.dtype) if p.requires_grad else None)
        return tuple(out)


def heads_with_grad(decoder, sequence_output: torch.Tensor, dropout=None) -> List[torch.Tensor]:
    params = [p for _, p in _param_items(decoder)]
    return list(_DecoderHeads.apply(decoder, sequence_output, dropout, *params))


class _DecoderHeadsLoss(torch.autograd.Function):
    """Heads + loss as ONE autograd node (fused tcgen05 configuration, OHEM off): forward = K1 + K2 with the loss
    reduced in K2's epilogue; backward = the backward tiles computing d loss / d logits in registers (no [B, P, C]
    gradient tensors, no separate loss kernels).  Outputs: out6 (five sub-losses + their ratio-weighted sum) and the
    five logits tensors.  A gradient arriving on the logits themselves (a caller hanging its own loss on them) takes
    the unfused route: explicit d loss / d logits + ``peneo_heads_bwd``."""

    @staticmethod
    def forward(ctx, decoder, x, drop, class_w, ratios, t0, t1, t2, t3, t4, *params):
        pack = decoder._weight_pack(x.device)
        ctx.in_dropout = decoder._input_dropout(drop)
        logits, out6, lctx = ops.heads_loss_forward(pack, x.detach(), [t0, t1, t2, t3, t4], class_w, ratios, drop, ctx.in_dropout,
                                                    save=True)
        ctx.decoder, ctx.dropout, ctx.lctx, ctx.logits = decoder, drop, lctx, logits
        ctx.save_for_backward(x)
        ctx.need_dx = x.requires_grad
        ctx.set_materialize_grads(False)
        return (out6, *logits)

    @staticmethod
    def backward(ctx, g6, *glogits):
        (x,) = ctx.saved_tensors
        dec = ctx.decoder
        if g6 is None:
            g6 = torch.zeros(6, dtype=torch.float32, device=x.device)
        if all(g is None for g in glogits):
            grads, dx = heads_backward(dec, x, None, ctx.need_dx, ctx.dropout, fused=(ctx.logits, ctx.lctx, g6),
                                       in_dropout=ctx.in_dropout)
            ctx.lctx = ctx.lctx[:4] + (None,)  # the saved activations were consumed (h now holds G)
        else:
            lws, tg, w3, r5, _ = ctx.lctx
            b, n = x.shape[0], x.shape[1]
            dl = ops.pair_loss_backward((lws, ctx.logits, tg, w3, r5, b, n), g6)
            dl = [d if g is None else d + g.float() for d, g in zip(dl, glogits)]
            grads, dx = heads_backward(dec, x, dl, ctx.need_dx, ctx.dropout, in_dropout=ctx.in_dropout)
        out = [None, dx.to(x.dtype) if dx is not None else None] + [None] * 8
        for k, p in _param_items(dec):
            out.append(grads[k].to(p.dtype) if p.requires_grad else None)
        return tuple(out)


def heads_loss_with_grad(decoder, sequence_output: torch.Tensor, tags, class_w, ratios, dropout=None):
    """-> (total loss, [five sub-losses], [five logits])"""
    params = [p for _, p in _param_items(decoder)]
    out = _DecoderHeadsLoss.apply(decoder, sequence_output, dropout, list(class_w), None if ratios is None else list(ratios),
                                  *tags, *params)
    out6, logits = out[0], list(out[1:])
    return out6[5], [out6[h] for h in range(5)], logits

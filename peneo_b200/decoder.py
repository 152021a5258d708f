"""``PEneoDecoderB200`` — drop-in for ``model.peneo_decoder.PEneoDecoder`` of ZeningLin/PEneo.

Same constructor, ``forward`` signature, return types, parameter names and state-dict keys
(model/peneo_decoder.py:201-443; checkpoint keys listed in SURVEY.md §5), so
``PEneoModel.from_pretrained`` / ``save_model`` / the trainer's ``"peneo_decoder"`` parameter
grouping keep working.  The arithmetic runs in libpeneo_b200.so:

* per-token projections (shrink MLP + the per-token halves of ``combine_fc``)   -> K1
* pair scoring + the five classifier heads, never materialising the pair tensor -> K2
* class-weighted pairwise loss                                                  -> loss kernels

The ``nn.Linear`` sub-modules below are parameter containers only; they are never called.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from ._lib import PREC_BF16, PREC_FP32
from .ops import HEAD_NAMES, DecoderDims, WeightPack

try:  # the reference subclasses transformers' ModelOutput; keep that when transformers is present
    from transformers.modeling_outputs import ModelOutput as _OutputBase
except Exception:  # pragma: no cover
    _OutputBase = object


@dataclass
class PEneoOutput(_OutputBase):
    """Field-for-field mirror of model/peneo_decoder.py:180-198."""

    loss: Optional[torch.Tensor] = None

    line_extraction_loss: Optional[torch.Tensor] = None
    ent_linking_h2h_loss: Optional[torch.Tensor] = None
    ent_linking_t2t_loss: Optional[torch.Tensor] = None
    line_grouping_h2h_loss: Optional[torch.Tensor] = None
    line_grouping_t2t_loss: Optional[torch.Tensor] = None

    line_extraction_shaking_outputs: Optional[torch.Tensor] = None
    ent_linking_h2h_shaking_outputs: Optional[torch.Tensor] = None
    ent_linking_t2t_shaking_outputs: Optional[torch.Tensor] = None
    line_grouping_h2h_shaking_outputs: Optional[torch.Tensor] = None
    line_grouping_t2t_shaking_outputs: Optional[torch.Tensor] = None

    attentions: Optional[Tuple[torch.FloatTensor]] = None
    hidden_states: Optional[Tuple[torch.FloatTensor]] = None
    orig_bbox: Optional[torch.Tensor] = None


class _ClassWeightedCE(nn.Module):
    """Holds the ``weight`` buffer and OHEM knobs of the reference's CrossEntropyLossOHEM
    (model/custom_loss.py:135-187) so that ``link_loss.weight`` / ``le_loss.weight`` stay in the
    state dict."""

    def __init__(self, weight: torch.Tensor, num_hard_positive: int, num_hard_negative: int):
        super().__init__()
        self.register_buffer("weight", weight)
        self.num_hard_positive = num_hard_positive
        self.num_hard_negative = num_hard_negative


def _cfg(config, name, default=None):
    return getattr(config, name, default) if not isinstance(config, dict) else config.get(name, default)


class PEneoDecoderB200(nn.Module):
    """PEneo pair-extraction head on B200 kernels.

    Extra, optional configuration (attributes on ``config``; absent = default):
      ``peneo_b200_precision``: ``"bf16"`` (tcgen05, default when the configuration allows it:
      shrink, hidden 768 -> d 384, 2 classifier layers) or ``"fp32"`` (CUDA-core fp32).
    """

    _warned_bf16_training = False

    def __init__(self, config, input_size: int) -> None:
        super().__init__()
        self.decoder_shrink = bool(_cfg(config, "peneo_decoder_shrink", True))
        backbone_config = _cfg(config, "backbone_config")
        hidden = backbone_config["hidden_size"]
        self.dropout_prob = backbone_config["hidden_dropout_prob"]
        self.num_layers = int(_cfg(config, "peneo_classifier_num_layers", 2))
        if self.decoder_shrink:
            d = hidden // 2
            self.shrink_projection = nn.ModuleDict({"0": nn.Linear(input_size, hidden), "3": nn.Linear(hidden, d)})
        else:
            d = input_size
        self.handshaking_kernel = nn.ModuleDict({"combine_fc": nn.Linear(2 * d, d)})
        self.inference_mode = bool(_cfg(config, "inference_mode", False))

        for name, classes in zip(HEAD_NAMES, ops.HEAD_CLASSES):
            if self.num_layers == 1:
                fc = nn.Linear(d, classes)
            else:
                layers = {str(3 * l): nn.Linear(d, d) for l in range(self.num_layers - 1)}
                layers[str(3 * (self.num_layers - 1))] = nn.Linear(d, classes)
                fc = nn.ModuleDict(layers)
            setattr(self, f"{name}_fc", fc)

        self.loss_ratio = _cfg(config, "peneo_loss_ratio", None)
        if self.loss_ratio is not None:
            assert len(self.loss_ratio) == 5, "loss_ratio must be a list of 5 elements"
        category_weights = _cfg(config, "peneo_category_weights", None)
        if category_weights is not None:
            assert len(category_weights) == 3, "category_weights must be a list of 3 elements"
            link_w = torch.tensor(category_weights).float()
            le_w = torch.tensor(category_weights[:-1]).float()
        # (the reference raises NameError here when category_weights is None: weights are mandatory)
        pos, neg = _cfg(config, "peneo_ohem_num_positive", -1), _cfg(config, "peneo_ohem_num_negative", -1)
        self.link_loss = _ClassWeightedCE(link_w, pos, neg)
        self.le_loss = _ClassWeightedCE(le_w, pos, neg)

        self.dims = DecoderDims(input_size, hidden if self.decoder_shrink else 0, d, self.decoder_shrink, self.num_layers)
        prec = _cfg(config, "peneo_b200_precision", None)
        if prec is None:
            # tensor cores wherever the widths allow it: the fused tcgen05 kernels for the shipped configuration, the
            # unfused tensor-core forward + kind::tf32 backward for the others
            prec = "bf16" if self.dims.bf16_forward_capable() else "fp32"
        self.set_precision(prec)
        self._pack: Optional[WeightPack] = None
        self._pack_key = None
        self._pack_readers = {}
        self._explicit_precision = _cfg(config, "peneo_b200_precision", None) is not None

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_reference(cls, ref_decoder: nn.Module, config, input_size: int) -> "PEneoDecoderB200":
        """Build from an instantiated reference ``PEneoDecoder`` (same weights, same device)."""
        new = cls(config, input_size)
        new.load_state_dict(ref_decoder.state_dict())
        p = next(ref_decoder.parameters())
        return new.to(p.device)

    def set_precision(self, prec: str) -> None:
        if prec not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        if prec == "bf16" and not self.dims.bf16_forward_capable():
            raise ValueError("the bf16 tensor-core paths need input / hidden / pair widths in multiples of 64; "
                             "use precision='fp32' for this configuration")
        self.precision = prec
        self._pack = None

    # ------------------------------------------------------------------ weights
    def _weight_pack(self, device) -> WeightPack:
        """Kernel-layout weights, re-packed when a parameter changed.  The pack is shared by every CUDA stream that
        runs the decoder (the caller's stream, HeadsDecodePipeline's compute stream): a stream other than the one
        that packed waits for the pack kernels, and a re-pack waits for the readers other streams registered with
        :meth:`_pack_read_done`."""
        state = {k: v for k, v in self.named_parameters()}
        key = (self.precision, str(device), tuple((p.data_ptr(), p._version) for p in state.values()))
        cur = torch.cuda.current_stream(device)
        if self._pack is None or self._pack_key != key:
            prec = PREC_BF16 if self.precision == "bf16" else PREC_FP32
            if self._pack is None or self._pack.prec != prec or self._pack.buf.device != device:
                self._pack = WeightPack(self.dims, prec, device)
                self._pack_readers = {}
            for ev in self._pack_readers.values():  # kernels of other streams still reading the old contents
                cur.wait_event(ev)
            self._pack_readers = {}
            self._pack.update({k: v.detach() for k, v in state.items()})
            self._pack_key = key
            self._pack_ready = torch.cuda.Event()
            self._pack_ready.record(cur)
            self._pack_stream = cur
        elif cur != self._pack_stream:
            cur.wait_event(self._pack_ready)
        return self._pack

    def _pack_read_done(self, device) -> None:
        """Called by multi-stream users after enqueueing kernels that read the pack on a side stream."""
        cur = torch.cuda.current_stream(device)
        if self._pack is not None and cur != getattr(self, "_pack_stream", cur):
            ev = torch.cuda.Event()
            ev.record(cur)
            self._pack_readers[cur.cuda_stream] = ev

    def _input_dropout(self, drop):
        """(p, seed) of the seam dropout for this step (same seed as the decoder's own dropout, its own mask site)."""
        p = float(getattr(self, "input_dropout_prob", 0.0))
        return (p, drop[1]) if (drop is not None and p > 0.0 and self.training) else None

    def _class_weights_host(self):
        """The three class weights as Python floats, cached (``.tolist()`` is a device sync)."""
        w = self.link_loss.weight
        key = (w.data_ptr(), w._version)
        if getattr(self, "_cw_key", None) != key:
            self._cw_host, self._cw_key = [float(v) for v in w.tolist()], key
        return self._cw_host

    def _fp32_pack(self, device) -> WeightPack:
        """fp32 kernel-layout weights (the backward pass always runs in fp32)."""
        if self.precision == "fp32":
            return self._weight_pack(device)
        state = {k: v for k, v in self.named_parameters()}
        key = (str(device), tuple((p.data_ptr(), p._version) for p in state.values()))
        if getattr(self, "_pack32", None) is None or self._pack32_key != key or self._pack32.buf.device != device:
            if getattr(self, "_pack32", None) is None or self._pack32.buf.device != device:
                self._pack32 = WeightPack(self.dims, PREC_FP32, device)
            self._pack32.update({k: v.detach() for k, v in state.items()})
            self._pack32_key = key
        return self._pack32

    # ------------------------------------------------------------------ forward
    def forward(
        self,
        sequence_output: torch.Tensor,
        orig_bbox: torch.Tensor = None,
        line_extraction_shaking_tag=None,
        ent_linking_head_rel_shaking_tag=None,
        ent_linking_tail_rel_shaking_tag=None,
        line_grouping_head_rel_shaking_tag=None,
        line_grouping_tail_rel_shaking_tag=None,
        **kwargs,
    ):
        if not sequence_output.is_cuda:
            raise RuntimeError("PEneoDecoderB200 runs on CUDA (sm_100) only; there is no CPU fallback")
        needs_grad = torch.is_grad_enabled() and (
            sequence_output.requires_grad or any(p.requires_grad for p in self.parameters())
        )
        # The decoder's own nn.Dropout modules (model/peneo_decoder.py:218, 221, 261) are active in train() mode
        # only.  The per-step seed comes from torch's generator, so torch.manual_seed() makes a step reproducible;
        # ``dropout_seed`` (attribute) pins it for tests.
        # ``input_dropout_prob`` (attribute, default 0): the nn.Dropout PEneoModel applies to the backbone output right
        # before the decoder (model/modeling_peneo.py:165), fused into the seam's gather pass; to use it replace
        # ``model.dropout`` by ``nn.Identity()`` and set this attribute to its probability.
        drop = None
        if self.training and (self.dropout_prob > 0 or getattr(self, "input_dropout_prob", 0.0) > 0):
            seed = getattr(self, "dropout_seed", None)
            drop = (float(self.dropout_prob), int(torch.randint(0, 2**62, (1,)).item()) if seed is None else int(seed))
        tags = [line_extraction_shaking_tag, ent_linking_head_rel_shaking_tag, ent_linking_tail_rel_shaking_tag,
                line_grouping_head_rel_shaking_tag, line_grouping_tail_rel_shaking_tag]
        ohem = (self.link_loss.num_hard_positive, self.link_loss.num_hard_negative)
        fused_out = None
        if needs_grad and not self.inference_mode:
            if self.precision == "bf16" and not self._explicit_precision and not PEneoDecoderB200._warned_bf16_training:
                PEneoDecoderB200._warned_bf16_training = True
                import warnings

                warnings.warn(
                    "PEneoDecoderB200 trains in its bf16 tensor-core mode by default (bf16 operands, fp32 accumulation, "
                    "tanh.approx SiLU): parameter gradients agree with the fp32 reference to ~1e-2 relative per tensor, "
                    "like the reference's own --fp16 AMP recipe.  Set config.peneo_b200_precision = 'fp32' for "
                    "fp32-exact (1e-6) training numerics.", stacklevel=2)
            fusable = (self.precision == "bf16" and self.dims.bf16_capable() and ohem == (-1, -1)
                       and getattr(self, "backward_precision", "bf16") == "bf16" and getattr(self, "fused_loss", True)
                       and all(torch.is_tensor(t) and t.is_cuda for t in tags))
            if fusable:
                # one autograd node for heads + loss: the loss is reduced in K2's epilogue and its backward happens
                # inside the backward tiles (model/peneo_decoder.py:355-428 as one sweep each way)
                shape = (sequence_output.shape[0], ops.shaking_len(sequence_output.shape[1]))
                for tg in tags:
                    assert tuple(tg.shape) == shape, "invalid input shape"
                from .train import heads_loss_with_grad

                total, subs, logits = heads_loss_with_grad(self, sequence_output, tags, self._class_weights_host(),
                                                           self.loss_ratio, drop)
                fused_out = (total, subs)
            else:
                from .autograd import decoder_forward_with_grad

                logits = decoder_forward_with_grad(self, sequence_output, drop)
        else:
            pack = self._weight_pack(sequence_output.device)
            logits = ops.heads_forward(pack, sequence_output.detach(), drop, self._input_dropout(drop))
        le, elh, elt, lgh, lgt = logits
        if self.inference_mode:
            return (le, elh, elt, lgh, lgt, orig_bbox)

        for lg, tg in zip(logits, tags):
            assert len(lg.shape) == len(tg.shape) + 1, "invalid input shape"
            assert lg.shape[:-1] == tg.shape
        if fused_out is not None:
            total, subs = fused_out
        elif ohem != (-1, -1):
            from .ohem import ohem_losses

            subs = ohem_losses(logits, tags, self.link_loss.weight, ohem)
            ratios = [1.0] * 5 if self.loss_ratio is None else self.loss_ratio
            total = sum(r * s for r, s in zip(ratios, subs))
        else:
            from .autograd import pair_loss_op

            total, subs = pair_loss_op(logits, tags, self._class_weights_host(), self.loss_ratio)
        return PEneoOutput(
            loss=total,
            line_extraction_loss=subs[0],
            ent_linking_h2h_loss=subs[1],
            ent_linking_t2t_loss=subs[2],
            line_grouping_h2h_loss=subs[3],
            line_grouping_t2t_loss=subs[4],
            line_extraction_shaking_outputs=le,
            ent_linking_h2h_shaking_outputs=elh,
            ent_linking_t2t_shaking_outputs=elt,
            line_grouping_h2h_shaking_outputs=lgh,
            line_grouping_t2t_shaking_outputs=lgt,
            orig_bbox=orig_bbox,
        )

"""peneo_b200 — B200-native (sm_100a) implementation of PEneo's decoder hot path.

Public surface mirrors the reference (ZeningLin/PEneo):
  PEneoDecoderB200 / PEneoOutput      <- model/peneo_decoder.py  PEneoDecoder / PEneoOutput
  HandshakingTaggingScheme            <- model/peneo_decoder.py  HandshakingTaggingScheme
  decode_peneo / sample_decode_peneo / parse_matrix_spots  <- pipeline/decode.py
  CrossEntropyLossOHEM                <- model/custom_loss.py    CrossEntropyLossOHEM
  evaluation.calculate_*_KVPE_metric  <- pipeline/evaluation.py
  eval_loop.prediction_loop / StreamingEvaluator  <- pipeline/trainer.py  PEneoTrainer.prediction_loop (per-batch decode)
"""
from .decode import decode_peneo, parse_matrix_spots, sample_decode_peneo  # noqa: F401
from .decoder import PEneoDecoderB200, PEneoOutput  # noqa: F401
from .eval_loop import StreamingEvaluator, prediction_loop  # noqa: F401
from .loss import CrossEntropyLossOHEM  # noqa: F401
from .pipeline import HeadsDecodePipeline  # noqa: F401
from .tagging import HandshakingTaggingScheme  # noqa: F401

__all__ = [
    "PEneoDecoderB200",
    "PEneoOutput",
    "HandshakingTaggingScheme",
    "HeadsDecodePipeline",
    "StreamingEvaluator",
    "prediction_loop",
    "CrossEntropyLossOHEM",
    "decode_peneo",
    "sample_decode_peneo",
    "parse_matrix_spots",
]

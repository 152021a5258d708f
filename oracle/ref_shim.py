"""TEST INFRASTRUCTURE ONLY — loader for the *real* reference hot path.

Imports ZeningLin/PEneo's decoder / loss / decode modules from ``/root/reference`` without
triggering the package ``__init__`` files (which pull in timm / detectron2 / an old
transformers API).  Only usable in the build container: the GPU box has
no ``/root/reference``.  Used by ``oracle/make_golden.py`` (fixture generation) and by the
``not gpu`` tests that pin the oracle against the reference when the reference is present.

Nothing under ``peneo_b200/`` may import this module.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PENEO_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "peneo_decoder.py"))


def load_reference():
    """Return a namespace with the reference's hot-path symbols."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # /root/reference is read-only
    for pkg in ("model", "data", "pipeline"):
        if pkg not in sys.modules or not hasattr(sys.modules[pkg], "__path__"):
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(REFERENCE_ROOT, pkg)]
            sys.modules[pkg] = m
    from model.configuration_peneo import PEneoConfig
    from model.peneo_decoder import (
        PEneoDecoder,
        HandshakingKernel,
        HandshakingTaggingScheme,
        PEneoOutput,
    )
    from model.custom_loss import CrossEntropyLossOHEM
    from pipeline.decode import decode_peneo, sample_decode_peneo, parse_matrix_spots
    from data.data_utils import merge_bbox
    from pipeline.evaluation import calculate_KVPE_metric, calculate_detail_KVPE_metric

    ns = types.SimpleNamespace(
        PEneoConfig=PEneoConfig,
        PEneoDecoder=PEneoDecoder,
        HandshakingKernel=HandshakingKernel,
        HandshakingTaggingScheme=HandshakingTaggingScheme,
        PEneoOutput=PEneoOutput,
        CrossEntropyLossOHEM=CrossEntropyLossOHEM,
        decode_peneo=decode_peneo,
        sample_decode_peneo=sample_decode_peneo,
        parse_matrix_spots=parse_matrix_spots,
        merge_bbox=merge_bbox,
        calculate_KVPE_metric=calculate_KVPE_metric,
        calculate_detail_KVPE_metric=calculate_detail_KVPE_metric,
    )
    return ns

"""TEST INFRASTRUCTURE ONLY — loader for the *real* reference hot path.

Imports ZeningLin/PEneo's decoder / loss / decode modules from ``/root/reference`` without
triggering the package ``__init__`` files (which pull in timm / detectron2 / an old
transformers API).  In the build container the modules come from the reference's
source tree; on the GPU box (no ``/root/reference``) from the byte code ``oracle/build_ref.py`` compiled
from that tree into the git-ignored ``oracle/_ref/``.  Used by ``oracle/make_golden.py`` (fixture
generation), by the tests that pin the oracle against the reference, and by ``bench.py``'s reference arm /
``cpu_baseline`` leg (which time the reference itself on the host cores).

Nothing under ``peneo_b200/`` may import this module.
"""
import importlib.abc
import importlib.util
import marshal
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PENEO_REFERENCE_ROOT", "/root/reference")
# byte code of the same modules, written by oracle/build_ref.py (git-ignored; travels to the GPU box)
COMPILED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def source_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "peneo_decoder.py"))


def compiled_available() -> bool:
    ver = os.path.join(COMPILED_ROOT, "PYTHON_VERSION")
    if not os.path.isfile(os.path.join(COMPILED_ROOT, "model", "peneo_decoder.refbc")) or not os.path.isfile(ver):
        return False
    with open(ver) as f:
        return f.read().strip() == "%d.%d" % sys.version_info[:2]


def reference_available() -> bool:
    return source_available() or compiled_available()


def reference_kind() -> str:
    """Where load_reference() takes the reference from: its source tree, or the byte code compiled from it."""
    return "source" if source_available() else "compiled" if compiled_available() else "none"


class _CompiledFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Imports ``<pkg>.<mod>`` from ``oracle/_ref/<pkg>/<mod>.refbc`` (a .pyc image written by build_ref.py)."""

    def find_spec(self, fullname, path=None, target=None):
        parts = fullname.split(".")
        if len(parts) != 2 or parts[0] not in ("model", "data", "pipeline"):
            return None
        fn = os.path.join(COMPILED_ROOT, parts[0], parts[1] + ".refbc")
        if not os.path.isfile(fn):
            return None
        return importlib.util.spec_from_loader(fullname, self, origin=fn)

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        with open(module.__spec__.origin, "rb") as f:
            data = f.read()
        module.__file__ = module.__spec__.origin
        exec(marshal.loads(data[16:]), module.__dict__)  # 16-byte .pyc header, then the marshalled code object


def load_reference():
    """Return a namespace with the reference's hot-path symbols."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT} nor compiled under {COMPILED_ROOT}")
    compiled = not source_available()
    root = COMPILED_ROOT if compiled else REFERENCE_ROOT
    sys.dont_write_bytecode = True  # /root/reference is read-only
    if compiled and not any(isinstance(f, _CompiledFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _CompiledFinder())
    for pkg in ("model", "data", "pipeline"):
        if pkg not in sys.modules or not hasattr(sys.modules[pkg], "__path__"):
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(root, pkg)]
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

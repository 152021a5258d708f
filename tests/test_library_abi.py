"""CPU-side checks of the C-ABI library: it builds, loads and exports every symbol that
include/peneo_b200.h declares.  No compute calls (no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from peneo_b200 import _lib, build

    build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    from peneo_b200 import _lib

    header = open(os.path.join(ROOT, "include", "peneo_b200.h")).read()
    declared = set(re.findall(r"^(?:int|size_t|const char\*)\s+(peneo_[a-z0-9_]+)\s*\(", header, re.M))
    assert declared, "no declarations found"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_abi_version_and_sizes(lib):
    from peneo_b200 import _lib

    assert lib.peneo_abi_version() == 4
    dims = _lib.Dims(768, 768, 384, 1, 2)
    assert lib.peneo_pack_bytes(dims, _lib.PREC_BF16) > 5 * 384 * 384 * 2
    assert lib.peneo_pack_bytes(dims, _lib.PREC_FP32) > 1_900_000 * 4
    # unsupported bf16 configuration is refused, with a message, not silently downgraded
    assert lib.peneo_pack_bytes(_lib.Dims(48, 0, 48, 0, 1), _lib.PREC_BF16) == 0
    assert b"PENEO_PREC_BF16" in lib.peneo_last_error()
    assert lib.peneo_decode_resolve_doc_ints(512, 1024) == 16 + 6 * 512 + 8 * 1024
    # widths in multiples of 64 outside the fused configuration: bf16 packs exist (unfused tensor-core forward) but the
    # mode has no backward, which the workspace query reports as 0
    for dims in (_lib.Dims(768, 0, 768, 0, 2), _lib.Dims(768, 768, 384, 1, 1), _lib.Dims(128, 128, 64, 1, 3)):
        nb = lib.peneo_pack_bytes(dims, _lib.PREC_BF16)
        assert nb > 5 * dims.d * 32 * 2, (dims.hin, dims.d, dims.num_layers, nb)
        assert lib.peneo_heads_bwd_workspace_bytes(dims, _lib.PREC_BF16, 2, 37) == 0
        assert lib.peneo_heads_bwd_workspace_bytes(dims, _lib.PREC_FP32, 2, 37) > 0
    assert lib.peneo_heads_bwd_workspace_bytes(_lib.Dims(768, 768, 384, 1, 2), _lib.PREC_BF16, 2, 37) > 0


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "peneo_b200")
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            src = open(os.path.join(pkg, name)).read()
            assert "peneo_oracle" not in src and "ref_shim" not in src and "import oracle" not in src, name


def test_decoder_refuses_cpu_tensors():
    import torch

    from peneo_b200 import PEneoDecoderB200

    class Cfg:
        backbone_config = {"hidden_size": 64, "hidden_dropout_prob": 0.1}
        peneo_decoder_shrink = True
        peneo_classifier_num_layers = 2
        peneo_loss_ratio = [1.0] * 5
        peneo_category_weights = [1.0, 10.0, 10.0]
        peneo_ohem_num_positive = -1
        peneo_ohem_num_negative = -1
        inference_mode = True

    dec = PEneoDecoderB200(Cfg, 64)
    assert dec.precision == "fp32"  # bf16 tcgen05 path only for the shipped 768/384/2-layer shape
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dec(torch.zeros(1, 4, 64))


def test_state_dict_keys_match_reference_checkpoint_interface():
    from peneo_b200 import PEneoDecoderB200, synth

    for shrink, L, hin in ((True, 2, 768), (False, 1, 48), (True, 3, 64)):
        class Cfg:
            backbone_config = {"hidden_size": hin, "hidden_dropout_prob": 0.1}
            peneo_decoder_shrink = shrink
            peneo_classifier_num_layers = L
            peneo_loss_ratio = [1.0] * 5
            peneo_category_weights = [1.0, 10.0, 10.0]
            peneo_ohem_num_positive = -1
            peneo_ohem_num_negative = -1
            inference_mode = False

        dec = PEneoDecoderB200(Cfg, hin)
        sd = synth.init_decoder_state(hin, hin, shrink, L)
        assert set(dec.state_dict().keys()) == set(sd.keys())
        for k, v in dec.state_dict().items():
            assert tuple(v.shape) == tuple(sd[k].shape), k


@pytest.mark.skipif(not __import__("ref_shim").reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("hin,hidden,shrink,L", [(768, 768, True, 2), (960, 768, True, 2), (48, 48, False, 1), (64, 64, True, 3)])
def test_module_is_checkpoint_compatible_with_the_live_reference(hin, hidden, shrink, L):
    """Parameter / buffer names and shapes of PEneoDecoderB200 equal those of the real PEneoDecoder, both ways
    (so `from_pretrained`, `save_model` and the trainer's "peneo_decoder" parameter grouping keep working)."""
    import ref_shim
    from peneo_b200 import PEneoDecoderB200

    ns = ref_shim.load_reference()
    cfg = ns.PEneoConfig(backbone_name="x", backbone_config={"hidden_size": hidden, "hidden_dropout_prob": 0.1},
                         peneo_decoder_shrink=shrink, peneo_classifier_num_layers=L, peneo_category_weights=[1, 10, 10])
    ref = ns.PEneoDecoder(cfg, hin)
    mine = PEneoDecoderB200(cfg, hin)
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
    assert [n for n, _ in ref.named_parameters()] == [n for n, _ in mine.named_parameters()]
    missing = mine.load_state_dict(a)
    assert not missing.missing_keys and not missing.unexpected_keys
    missing = ref.load_state_dict(b)
    assert not missing.missing_keys and not missing.unexpected_keys

"""Pins the CPU oracle (oracle/peneo_oracle.py) against outputs of the real reference
(tests/golden/*.pt, written by oracle/make_golden.py).  CPU only."""
import torch
import pytest

import peneo_oracle as orc
from peneo_b200 import synth


def _state(case):
    if case["state_dict"] is not None:
        return case["state_dict"]
    return synth.init_decoder_state(case["hin"], case["hidden"], case["shrink"], case["num_layers"],
                                    seed=case["seed"], trained_like=(case.get("init", "trained") == "trained"))


def _close(a, b, tol):
    scale = max(b.abs().max().item(), 1e-30)
    return (a - b).abs().max().item() <= tol * scale


def test_heads_ref_style_and_chunked_match_reference(golden):
    for case in golden("heads.pt"):
        sd = _state(case)
        x = synth.hidden_states(case["batch"], case["seq_len"], case["hin"], doc_id0=case["x_doc_id0"])
        p32 = orc.split_params(sd, torch.float32)
        p64 = orc.split_params(sd, torch.float64)
        out32 = orc.heads_ref_style(p32, x)
        out64 = orc.heads_chunked(p64, x.double(), row_block=4)
        for k in range(5):
            ref = case["logits"][k]
            assert out32[k].shape == ref.shape
            assert _close(out32[k], ref, 2e-5), (case["name"], k)
            assert _close(out64[k].float(), ref, 2e-5), (case["name"], k)


def test_loss_matches_reference_including_ohem_quirks(golden):
    for case in golden("loss.pt"):
        got = orc.ce_loss(case["logits"], case["target"], case["weight"], *case["ohem"])
        assert abs(got.item() - case["loss"].item()) <= 2e-6 * max(1.0, abs(case["loss"].item())), case["name"]


def test_train_forward_loss_and_grads_match_reference_autograd(golden):
    for case in golden("train.pt"):
        sd = synth.init_decoder_state(case["hin"], case["hidden"], case["shrink"], case["num_layers"],
                                      seed=case["seed"], trained_like=True)
        x = synth.hidden_states(case["batch"], case["seq_len"], case["hin"], doc_id0=case["x_doc_id0"])
        docs = [synth.make_document(case["seq_len"], doc_id=case["doc_id0"] + b) for b in range(case["batch"])]
        tags = [torch.stack([d.tags()[k] for d in docs]) for k in range(5)]
        loss, subs, grads, dx = orc.loss_and_grads(sd, x, tags, [1.0, 10.0, 10.0], case["ratios"])
        assert abs(loss.item() - case["loss"].item()) < 1e-5 * max(1, abs(case["loss"].item()))
        for k in range(5):
            assert abs(subs[k].item() - case["sub_losses"][k].item()) < 1e-5 * max(1, abs(case["sub_losses"][k].item()))
        assert _close(dx.float(), case["dx"], 2e-4)
        for key, g in case["grads"].items():
            assert _close(grads[key].float(), g, 2e-4), (case["name"], key)


def _regen(case):
    r = case["regen"]
    doc = synth.make_document(r["n"], doc_id=r["doc_id"], style=r["style"])
    if r["kind"] == "planted_gt":
        return doc.tags()
    dt = torch.bfloat16 if r.get("dtype") == "bf16" else torch.float32
    return synth.planted_logits(doc, seed=r["doc_id"], dtype=dt)


def test_decode_matches_reference_objects_including_dict_order(golden):
    g = golden("decode.pt")
    for case in g["samples"]:
        sh = case["shakings"] if case["shakings"] is not None else _regen(case)
        for k in range(5):
            assert orc.get_spots(sh[k], case["seq_len"]) == case["spots"][k], (case["name"], k)
        bbox = case["bbox"]
        got = orc.sample_decode(case["text"], sh, case["seq_len"], bbox=bbox, decode_gt=case["decode_gt"],
                                score_thresh=case["score_thresh"])
        ref = case["result"]
        assert got[0] == ref[0] and got[1] == ref[1], case["name"]
        for a, b in zip(got[2:], ref[2:]):
            assert a == b and list(a.items()) == list(b.items()), case["name"]


def test_decode_batch_matches_reference(golden):
    b = golden("decode.pt")["batch"]
    docs = [synth.make_document(b["n"], doc_id=d) for d in b["doc_ids"]]
    outs = [[synth.planted_logits(d, seed=did)[k] for d, did in zip(docs, b["doc_ids"])] for k in range(5)]
    tg = [[d.tags()[k] for d in docs] for k in range(5)]
    got = orc.decode_batch([d.text for d in docs], outs, tg, [d.bbox for d in docs], ["f0", "f1", "f2"])
    ref = b["result"]
    assert got[2] == ref[2]
    for side in (0, 1):
        for s_got, s_ref in zip(got[side], ref[side]):
            assert s_got[0] == s_ref[0] and s_got[1] == s_ref[1]
            for a, c in zip(s_got[2:], s_ref[2:]):
                assert list(a.items()) == list(c.items())


def test_parse_matrix_spots_known_answers(golden):
    for kat in golden("decode.pt")["parse_kats"]:
        got = orc.parse_matrix_spots(kat["spots"], kat["top"], kat["triu"], kat["thresh"])
        assert list(got.items()) == list(kat["result"].items())
    # SURVEY.md appendix A.5
    spots = [(0, 5, 1, 0.9), (0, 6, 1, 0.9), (1, 5, 1, 0.95), (2, 7, 2, 0.8)]
    assert orc.parse_matrix_spots(spots, True, True) == {1: 5, 7: 2}
    assert orc.parse_matrix_spots(spots, False, True) == {0: [5, 6], 1: [5], 7: [2]}


def test_tag_codec_matches_reference(golden):
    for case in golden("tags.pt"):
        assert torch.equal(orc.spots_to_tags(case["batch_spots"], case["n"]), case["tags"])


def test_index_formula_roundtrip():
    import numpy as np

    for n in (1, 2, 7, 64, 511, 2048):
        p = np.arange(orc.shaking_len(n))
        i, j = orc.unflatten_index(p, n)
        assert (i <= j).all() and (j < n).all()
        assert (i * n - i * (i - 1) // 2 + (j - i) == p).all()


@pytest.mark.skipif(not __import__("ref_shim").reference_available(), reason="reference tree not present")
def test_oracle_against_live_reference_larger_case():
    import ref_shim

    ns = ref_shim.load_reference()
    cfg = ns.PEneoConfig(backbone_name="x", backbone_config={"hidden_size": 768, "hidden_dropout_prob": 0.1},
                         peneo_category_weights=[1, 10, 10], inference_mode=True)
    dec = ns.PEneoDecoder(cfg, 768).eval()
    sd = synth.init_decoder_state(seed=1, trained_like=True)
    dec.load_state_dict(sd)
    x = synth.hidden_states(1, 40, 768)
    with torch.no_grad():
        ref = dec(x)
    got = orc.heads_chunked(orc.split_params(sd, torch.float64), x.double())
    for k in range(5):
        assert _close(got[k].float(), ref[k], 2e-5)


@pytest.mark.skipif(not __import__("ref_shim").reference_available(), reason="reference tree not present")
def test_tag_codec_edge_cases_against_live_reference():
    """Python-indexing semantics of the reference's N x N lookup table (model/peneo_decoder.py:55-59, 70): spots
    below the diagonal hit cell 0, negative indices count from the end, out-of-range indices raise IndexError."""
    import ref_shim

    ns = ref_shim.load_reference()
    edge = [[(3, 1, 2), (-1, -1, 1), (1, 2, 1)], [(0, -2, 2), (4, 0, 1)]]
    ref = ns.HandshakingTaggingScheme.spots2shaking_tag4batch(edge, seq_len=5)
    assert torch.equal(orc.spots_to_tags(edge, 5), ref)
    for bad in ([[(0, 5, 1)]], [[(-6, 0, 1)]]):
        with pytest.raises(IndexError):
            ns.HandshakingTaggingScheme.spots2shaking_tag4batch(bad, seq_len=5)
        with pytest.raises(IndexError):
            orc.spots_to_tags(bad, 5)


def test_compiled_reference_recipe_round_trips(tmp_path, monkeypatch):
    """oracle/build_ref.py: byte code compiled from the reference tree imports without its sources and computes the
    same logits as the source tree (skipped where /root/reference does not exist, e.g. on the GPU box)."""
    import importlib
    import subprocess
    import sys

    import build_ref
    import ref_shim

    if not ref_shim.source_available():
        pytest.skip("reference source tree not present")
    assert build_ref.build(quiet=True)
    code = (
        "import sys, torch; sys.path.insert(0, %r); import ref_shim; "
        "assert ref_shim.reference_kind() == 'compiled', ref_shim.reference_kind(); ns = ref_shim.load_reference(); "
        "cfg = ns.PEneoConfig(backbone_name='x', backbone_config={'hidden_size': 64, 'hidden_dropout_prob': 0.1}, "
        "peneo_category_weights=[1, 10, 10], inference_mode=True); torch.manual_seed(0); "
        "d = ns.PEneoDecoder(cfg, 64).eval(); print(float(d(torch.ones(1, 5, 64))[1].sum()))"
    ) % build_ref.HERE
    env = dict(__import__("os").environ, PENEO_REFERENCE_ROOT="/nonexistent")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    ns = ref_shim.load_reference()
    cfg = ns.PEneoConfig(backbone_name="x", backbone_config={"hidden_size": 64, "hidden_dropout_prob": 0.1},
                         peneo_category_weights=[1, 10, 10], inference_mode=True)
    torch.manual_seed(0)
    d = ns.PEneoDecoder(cfg, 64).eval()
    assert abs(float(d(torch.ones(1, 5, 64))[1].sum()) - float(out.stdout.strip())) < 1e-6

"""CPU tests of the decode host glue (peneo_b200/_hostglue, csrc/hostglue.c): record blocks in the layout
peneo_decode_resolve writes -> the reference's Python result objects.  The records are built here from
the oracle's own decode, so the comparison is end to end against pipeline/decode.py semantics."""
import numpy as np
import torch

import peneo_oracle as orc
from peneo_b200 import decode as dec
from peneo_b200 import synth


def _record_from_oracle(text, shakings, n, cap, decode_gt=False, score_thresh=0):
    """Oracle decode of one document + the int32 record block a correct K4 would have produced."""
    res = orc.sample_decode(text, shakings, n, decode_gt=decode_gt, score_thresh=score_thresh)
    _pairs, _lines, le, el_head, el_tail, lg_head, lg_tail = res
    rec = np.zeros(16 + 6 * n + 8 * cap, dtype=np.int32)

    def put(off, items):
        flat = [v for kv in items for v in kv]
        rec[off : off + len(flat)] = flat
        return len(items)

    off = 16
    rec[0] = put(off, list(le.items()))
    off += 2 * n
    rec[1] = put(off, list(lg_head.items()))
    off += 2 * n
    rec[2] = put(off, list(lg_tail.items()))
    off += 2 * n
    # EL lists in arrival order: re-derive from the spot lists (dict-of-lists loses interleaving)
    def edges(spots):
        out = []
        for i, j, tag, score in spots:
            if tag == 0 or score < score_thresh:
                continue
            out.append((j, i) if tag == 2 else (i, j))
        return out

    elh = edges(orc.get_spots(shakings[1], n))
    elt = edges(orc.get_spots(shakings[2], n))
    rec[3] = put(off, elh)
    off += 2 * cap
    rec[4] = put(off, elt)
    off += 2 * cap
    kv = []
    for kh, vh in elh:
        kt, vt = le.get(kh), le.get(vh)
        if kt is None or vt is None:
            continue
        ks, k_last = orc._walk(kh, kt, le, lg_head, lg_tail)
        vs, v_last = orc._walk(vh, vt, le, lg_head, lg_tail)
        tails = el_tail.get(k_last)
        if tails is not None and v_last in tails:
            kv.append((kh, vh, len(ks), len(vs)))
    rec[5] = put(off, kv)
    return res, rec


def _check(text, shakings, n, bbox=None, **kw):
    cap = max(8 * n, 1024)
    ref, rec = _record_from_oracle(text, shakings, n, cap, **kw)
    if bbox is not None:
        ref = orc.sample_decode(text, shakings, n, bbox=bbox, **kw)
    dd = dec.DeviceDecode(n, cap, 1, rec.reshape(1, -1), np.zeros((1, 5), np.int32), None)
    got = dec.assemble_many(dd, [0], [text], None if bbox is None else [bbox])[0]
    assert got[0] == ref[0] and got[1] == ref[1]
    for a, b in zip(got[2:], ref[2:]):
        assert a == b and list(a.items()) == list(b.items())
    return got


def test_planted_documents_with_and_without_boxes():
    for n, did, style in [(24, 1, "rfund"), (63, 2, "rfund"), (95, 3, "sibr")]:
        doc = synth.make_document(n, doc_id=did, style=style)
        logits = synth.planted_logits(doc, seed=did)
        got = _check(doc.text, logits, n)
        assert len(got[1]) > 0
        _check(doc.text, logits, n, bbox=torch.tensor(doc.bbox))
        _check(doc.text, doc.tags(), n, decode_gt=True)
        _check(doc.text[: doc.n_real], logits, n)  # text shorter than N: slicing tolerates the overrun


def test_dense_random_and_threshold():
    g = torch.Generator().manual_seed(5)
    for n in (6, 12, 17):
        sh = [torch.randn(n * (n + 1) // 2, c, generator=g) for c in (2, 3, 3, 3, 3)]
        text = [f"t{i}|" for i in range(n)]
        _check(text, sh, n)
        _check(text, sh, n, score_thresh=0.6)


def test_cycle_emits_1001_segments():
    n = 12
    tags = [torch.zeros(n * (n + 1) // 2, dtype=torch.int64) for _ in range(5)]
    for h, t in [(0, 1), (4, 5), (8, 9), (10, 11)]:
        tags[0][synth.shaking_index(h, t, n)] = 1
    for a, b in [(0, 4), (4, 8)]:
        tags[3][synth.shaking_index(a, b, n)] = 1
    tags[3][synth.shaking_index(0, 8, n)] = 2
    for a, b in [(1, 5), (5, 9)]:
        tags[4][synth.shaking_index(a, b, n)] = 1
    tags[4][synth.shaking_index(1, 9, n)] = 2
    tags[1][synth.shaking_index(0, 10, n)] = 1
    tags[2][synth.shaking_index(5, 11, n)] = 1
    got = _check([f"c{i}." for i in range(n)], tags, n, decode_gt=True)
    assert len(got[0]) == 1 and len(got[0][0][0]) == 1001 * 6  # "cX.cY." per segment


def test_many_documents_in_one_call_and_bad_input():
    docs = [synth.make_document(31, doc_id=40 + b) for b in range(3)]
    cap = 1024
    refs, recs = zip(*[_record_from_oracle(d.text, synth.planted_logits(d, seed=40 + b), 31, cap) for b, d in enumerate(docs)])
    dd = dec.DeviceDecode(31, cap, 3, np.stack(recs), np.zeros((3, 5), np.int32), None)
    got = dec.assemble_many(dd, [2, 0], [docs[2].text, docs[0].text])
    assert got[0][0] == refs[2][0] and got[1][1] == refs[0][1]
    import pytest

    with pytest.raises(IndexError):
        dec.assemble_many(dd, [3], [docs[0].text])
    bad = np.stack(recs).copy()
    bad[0, 0] = 10 ** 6
    with pytest.raises(ValueError):
        dec.assemble_many(dec.DeviceDecode(31, cap, 3, bad, np.zeros((3, 5), np.int32), None), [0], [docs[0].text])

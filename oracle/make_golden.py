"""TEST INFRASTRUCTURE ONLY — writes tests/golden/*.pt from the REAL reference.

Run in the build container (needs /root/reference, see oracle/ref_shim.py):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Every fixture holds inputs (or the seeds that regenerate them through peneo_b200.synth) and
the outputs the unmodified reference code produced for them.  The fixtures travel to the GPU
box; the reference does not.
"""
import os
import random
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_shim  # noqa: E402
from peneo_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
HEADS = ("line_extraction", "ent_linking_h2h", "ent_linking_t2t", "line_grouping_h2h", "line_grouping_t2t")


def build_ref_decoder(ns, hin, hidden, shrink, num_layers, ohem=(-1, -1), ratios=(1.0,) * 5, weights=(1.0, 10.0, 10.0),
                      inference_mode=False):
    cfg = ns.PEneoConfig(
        backbone_name="synthetic",
        backbone_config={"hidden_size": hidden, "hidden_dropout_prob": 0.1},
        peneo_decoder_shrink=shrink,
        peneo_classifier_num_layers=num_layers,
        peneo_loss_ratio=list(ratios),
        peneo_category_weights=list(weights),
        peneo_ohem_num_positive=ohem[0],
        peneo_ohem_num_negative=ohem[1],
        inference_mode=inference_mode,
    )
    return ns.PEneoDecoder(cfg, input_size=hin).eval()


def heads_cases(ns):
    cases = []
    specs = [
        # name, hin, hidden, shrink, L, B, N, init
        ("default_dims_hf_init", 768, 768, True, 2, 1, 9, "hf"),
        ("lilt_dims_hf_init", 960, 768, True, 2, 1, 7, "hf"),
        ("small_trained_like", 64, 64, True, 2, 2, 13, "trained"),
        ("small_noshrink_L1", 48, 48, False, 1, 2, 11, "trained"),
        ("small_L3", 64, 64, True, 3, 1, 10, "trained"),
        ("small_torch_default_init", 64, 64, True, 2, 2, 8, "torch"),
    ]
    for name, hin, hidden, shrink, L, B, N, init in specs:
        dec = build_ref_decoder(ns, hin, hidden, shrink, L, inference_mode=True)
        if init == "torch":
            torch.manual_seed(11)
            dec = build_ref_decoder(ns, hin, hidden, shrink, L, inference_mode=True)
            sd = {k: v.clone() for k, v in dec.state_dict().items()}
            sd_store = sd
            seed = None
        else:
            seed = 3 if init == "hf" else 5
            sd = synth.init_decoder_state(hin, hidden, shrink, L, seed=seed, trained_like=(init == "trained"))
            dec.load_state_dict(sd)
            sd_store = None  # regenerated from the seed
        x = synth.hidden_states(B, N, hin, doc_id0=17)
        with torch.no_grad():
            out = dec(x)
        cases.append(
            dict(name=name, hin=hin, hidden=hidden, shrink=shrink, num_layers=L, batch=B, seq_len=N, init=init,
                 seed=seed, state_dict=sd_store, x_doc_id0=17, logits=[o.clone() for o in out[:5]])
        )
    return cases


def loss_cases(ns):
    cases = []
    g = torch.Generator().manual_seed(99)
    for name, C, M, npos, ohem in [
        ("link_ohem_off", 3, 400, 25, (-1, -1)),
        ("le_ohem_off", 2, 300, 30, (-1, -1)),
        ("link_ohem_5_20", 3, 400, 25, (5, 20)),
        ("link_ohem_neg_only", 3, 400, 25, (-1, 7)),
        ("link_ohem_pos_only", 3, 400, 25, (6, -1)),
        ("link_ohem_more_than_available", 3, 60, 4, (50, 500)),
    ]:
        logits = torch.randn(M, C, generator=g) * 2.0
        target = torch.zeros(M, dtype=torch.int64)
        idx = torch.randperm(M, generator=g)[:npos]
        target[idx] = torch.randint(1, C, (npos,), generator=g)
        w = torch.tensor([1.0, 10.0, 10.0][:C])
        crit = ns.CrossEntropyLossOHEM(num_hard_positive=ohem[0], num_hard_negative=ohem[1], weight=w)
        val = crit(logits, target)
        cases.append(dict(name=name, logits=logits, target=target, weight=w, ohem=ohem, loss=val.clone()))
    return cases


def train_cases(ns):
    """Full decoder forward with tags -> PEneoOutput losses and autograd gradients."""
    cases = []
    for name, hin, hidden, shrink, L, B, N, ratios in [
        ("small_L2", 64, 64, True, 2, 2, 12, (1.0, 1.0, 1.0, 1.0, 1.0)),
        ("small_L2_ratios", 64, 64, True, 2, 1, 9, (1.0, 0.5, 2.0, 0.25, 1.5)),
        ("small_noshrink_L1", 48, 48, False, 1, 2, 10, (1.0, 1.0, 1.0, 1.0, 1.0)),
    ]:
        dec = build_ref_decoder(ns, hin, hidden, shrink, L, ratios=ratios)
        sd = synth.init_decoder_state(hin, hidden, shrink, L, seed=21, trained_like=True)
        dec.load_state_dict(sd)
        x = synth.hidden_states(B, N, hin, doc_id0=40).requires_grad_(True)
        docs = [synth.make_document(N, doc_id=100 + b) for b in range(B)]
        tags = [torch.stack([d.tags()[k] for d in docs]) for k in range(5)]
        out = dec(x, None, *[tags[k] for k in (0, 1, 2, 3, 4)])
        out.loss.backward()
        grads = {k: p.grad.clone() for k, p in dec.named_parameters()}
        cases.append(
            dict(name=name, hin=hin, hidden=hidden, shrink=shrink, num_layers=L, batch=B, seq_len=N, ratios=ratios,
                 seed=21, x_doc_id0=40, doc_id0=100,
                 loss=out.loss.detach().clone(),
                 sub_losses=[out.line_extraction_loss.detach().clone(), out.ent_linking_h2h_loss.detach().clone(),
                             out.ent_linking_t2t_loss.detach().clone(), out.line_grouping_h2h_loss.detach().clone(),
                             out.line_grouping_t2t_loss.detach().clone()],
                 logits=[out.line_extraction_shaking_outputs.detach().clone(),
                         out.ent_linking_h2h_shaking_outputs.detach().clone(),
                         out.ent_linking_t2t_shaking_outputs.detach().clone(),
                         out.line_grouping_h2h_shaking_outputs.detach().clone(),
                         out.line_grouping_t2t_shaking_outputs.detach().clone()],
                 grads=grads, dx=x.grad.clone())
        )
    return cases


def decode_cases(ns):
    tagger = ns.HandshakingTaggingScheme()
    cases = []

    def run(name, text, shakings, n, bbox=None, decode_gt=False, score_thresh=0, store_inputs=True, regen=None):
        res = ns.sample_decode_peneo(tagger, text, *shakings, bbox=bbox, seq_len=n, decode_gt=decode_gt,
                                     score_thresh=score_thresh)
        spots = [tagger.get_spots_from_shaking_tag(s, seq_len=n) for s in shakings]
        cases.append(dict(name=name, text=text, seq_len=n, bbox=bbox, decode_gt=decode_gt, score_thresh=score_thresh,
                          shakings=[s.clone() for s in shakings] if store_inputs else None, regen=regen,
                          result=res, spots=spots))

    # planted documents: prediction decode and GT decode
    for n, did, style in [(24, 1, "rfund"), (63, 2, "rfund"), (95, 3, "sibr"), (40, 4, "sibr")]:
        doc = synth.make_document(n, doc_id=did, style=style)
        logits = synth.planted_logits(doc, seed=did)
        run(f"planted_pred_n{n}", doc.text, logits, n, regen=dict(kind="planted", n=n, doc_id=did, style=style),
            store_inputs=False)
        run(f"planted_pred_bbox_n{n}", doc.text, logits, n, bbox=torch.tensor(doc.bbox),
            regen=dict(kind="planted", n=n, doc_id=did, style=style), store_inputs=False)
        run(f"planted_gt_n{n}", doc.text, doc.tags(), n, decode_gt=True,
            regen=dict(kind="planted_gt", n=n, doc_id=did, style=style), store_inputs=False)
    # text shorter than N (python slicing tolerates the overrun)
    doc = synth.make_document(31, doc_id=9)
    run("short_text", doc.text[: doc.n_real], synth.planted_logits(doc, seed=9), 31)
    # dense garbage logits: many competing spots, exercises top-1 / no-fallback / arrival order
    g = torch.Generator().manual_seed(5)
    for n in (6, 12, 17):
        sh = [torch.randn(n * (n + 1) // 2, c, generator=g) for c in (2, 3, 3, 3, 3)]
        text = [f"t{i}|" for i in range(n)]
        run(f"dense_random_n{n}", text, sh, n)
        run(f"dense_random_thresh_n{n}", text, sh, n, score_thresh=0.6)
    # exact score ties (quantised logits) -> first-arrival tie-break
    for n in (8, 14):
        sh = [torch.randint(-1, 2, (n * (n + 1) // 2, c), generator=g).float() * 2.0 for c in (2, 3, 3, 3, 3)]
        run(f"ties_n{n}", [f"w{i} " for i in range(n)], sh, n)
    # GT decode with duplicate heads (top_score_only=False then first element)
    n = 10
    tags = [torch.zeros(n * (n + 1) // 2, dtype=torch.int64) for _ in range(5)]
    for k, sp in enumerate([[(0, 2, 1), (0, 4, 1), (5, 7, 1), (3, 4, 1)], [(0, 5, 1), (3, 5, 2)], [(2, 7, 1), (4, 7, 2)],
                            [(0, 3, 1)], [(2, 4, 1)]]):
        for i, j, t in sp:
            tags[k][synth.shaking_index(i, j, n)] = t
    run("gt_duplicates", [f"g{i}" for i in range(n)], tags, n, decode_gt=True)
    # 3-cycle in line grouping: the 1000-step cap is observable
    n = 12
    tags = [torch.zeros(n * (n + 1) // 2, dtype=torch.int64) for _ in range(5)]
    le = [(0, 1), (4, 5), (8, 9), (10, 11)]
    for h, t in le:
        tags[0][synth.shaking_index(h, t, n)] = 1
    # heads 0->4->8->0 ; tails 1->5->9->1 ; (8,0) / (9,1) are lower-triangular -> stored flipped with tag 2
    for (a, b) in [(0, 4), (4, 8)]:
        tags[3][synth.shaking_index(a, b, n)] = 1
    tags[3][synth.shaking_index(0, 8, n)] = 2
    for (a, b) in [(1, 5), (5, 9)]:
        tags[4][synth.shaking_index(a, b, n)] = 1
    tags[4][synth.shaking_index(1, 9, n)] = 2
    tags[1][synth.shaking_index(0, 10, n)] = 1
    # after 1000 steps from head 0: 1000 % 3 = 1 -> current line is (4,5)
    tags[2][synth.shaking_index(5, 11, n)] = 1
    run("cycle_cap_gt", [f"c{i}." for i in range(n)], tags, n, decode_gt=True)
    logits = []
    for k, t in enumerate(tags):
        c = 2 if k == 0 else 3
        z = torch.zeros(t.shape[0], c)
        z[:, 0] = 4.0
        z[torch.arange(t.shape[0]), t] += 8.0 * (t > 0).float()
        logits.append(z)
    run("cycle_cap_pred", [f"c{i}." for i in range(n)], logits, n)
    # empty document (no spot anywhere)
    n = 5
    run("empty", ["a"] * n, [torch.zeros(n * (n + 1) // 2, dtype=torch.int64) for _ in range(5)], n, decode_gt=True)
    z = [torch.zeros(n * (n + 1) // 2, c) for c in (2, 3, 3, 3, 3)]
    for t in z:
        t[:, 0] = 1.0
    run("empty_pred", ["a"] * n, z, n)
    # bf16 logits: softmax runs in bf16, scores collapse to a few values
    doc = synth.make_document(47, doc_id=12)
    run("planted_bf16", doc.text, synth.planted_logits(doc, seed=12, dtype=torch.bfloat16), 47,
        regen=dict(kind="planted", n=47, doc_id=12, style="rfund", dtype="bf16"), store_inputs=False)

    # batch-level decode_peneo
    docs = [synth.make_document(29, doc_id=30 + b) for b in range(3)]
    outs = [[synth.planted_logits(d, seed=30 + b)[k] for b, d in enumerate(docs)] for k in range(5)]
    tg = [[d.tags()[k] for d in docs] for k in range(5)]
    res = ns.decode_peneo(tagger, [d.text for d in docs], *outs, *tg, [d.bbox for d in docs], ["f0", "f1", "f2"])
    batch_case = dict(n=29, doc_ids=[30, 31, 32], result=res)

    kats = []
    spots = [(0, 5, 1, 0.9), (0, 6, 1, 0.9), (1, 5, 1, 0.95), (2, 7, 2, 0.8)]
    for top in (True, False):
        for triu in (True, False):
            for th in (0, 0.85):
                kats.append(dict(spots=spots, top=top, triu=triu, thresh=th,
                                 result=ns.parse_matrix_spots(spots, top, triu, th)))
    rnd = random.Random(4)
    for _ in range(20):
        sp = [(rnd.randint(0, 6), rnd.randint(0, 6), rnd.randint(0, 2), rnd.choice([0.5, 0.7, 0.9, 0.9, 0.99]))
              for _ in range(rnd.randint(0, 14))]
        sp = [(min(a, b), max(a, b), t, s) for a, b, t, s in sp]
        for top in (True, False):
            kats.append(dict(spots=sp, top=top, triu=True, thresh=0,
                             result=ns.parse_matrix_spots(sp, top, True, 0)))
    return dict(samples=cases, batch=batch_case, parse_kats=kats)


def tag_cases(ns):
    tagger = ns.HandshakingTaggingScheme()
    cases = []
    for n, did in [(9, 1), (16, 2)]:
        docs = [synth.make_document(n, doc_id=did + b) for b in range(2)]
        for k in range(5):
            batch_spots = [d.spots[k] for d in docs]
            ref = tagger.spots2shaking_tag4batch(batch_spots, seq_len=n)
            cases.append(dict(n=n, batch_spots=batch_spots, tags=ref))
    return cases


def eval_cases(ns):
    """pipeline/evaluation.py on decoded synthetic documents: predictions = decode of planted logits of a
    *perturbed* document (so that TP / FP / FN all occur), ground truth = GT decode of the document."""
    tagger = ns.HandshakingTaggingScheme()
    preds, gts, names = [], [], []
    for did, n in [(1, 63), (2, 95), (3, 40), (4, 63), (5, 127)]:
        doc = synth.make_document(n, doc_id=did)
        other = synth.make_document(n, doc_id=did + (0 if did % 2 else 50))  # odd ids: perfect prediction
        pred = ns.sample_decode_peneo(tagger, doc.text, *synth.planted_logits(other, seed=did), seq_len=n)
        gt = ns.sample_decode_peneo(tagger, doc.text, *doc.tags(), seq_len=n, decode_gt=True)
        preds.append(pred), gts.append(gt), names.append(f"file_{did}.json")
    # a duplicated file name (distributed sampler padding) and an empty sample
    preds.append(preds[0]), gts.append(gts[0]), names.append(names[0])
    empty = ([], [], {}, {}, {}, {}, {})
    preds.append(empty), gts.append(empty), names.append("empty.json")
    kv = ns.calculate_KVPE_metric(preds, gts, names)
    detail = ns.calculate_detail_KVPE_metric(preds, gts, names)
    return dict(preds=preds, gts=gts, names=names, kvpe=kv, detail=detail)


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = ref_shim.load_reference()
    torch.manual_seed(0)
    torch.save(heads_cases(ns), os.path.join(OUT, "heads.pt"))
    torch.save(loss_cases(ns), os.path.join(OUT, "loss.pt"))
    torch.save(train_cases(ns), os.path.join(OUT, "train.pt"))
    torch.save(decode_cases(ns), os.path.join(OUT, "decode.pt"))
    torch.save(tag_cases(ns), os.path.join(OUT, "tags.pt"))
    if os.environ.get("PENEO_GOLDEN_ONLY", "") in ("", "eval"):
        torch.save(eval_cases(ns), os.path.join(OUT, "eval.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()

"""Synthetic RFUND / SIBR-shaped documents for parity tests and benchmarks.

The datasets themselves are not shipped with the reference (its .gitignore:17-18), so document
*shape* follows what ``data/datasets/sibr.py:261-408`` produces: lines of a few tokens, entities
made of 1-3 consecutive lines, line-extraction spots ``(start, end, 1)``, line-grouping
head-to-head / tail-to-tail spots between consecutive lines of an entity, entity-linking spots
from a question entity to the following answer entity, and flipped relations stored with tag 2
(sibr.py:315-347, 392-408).  Recipe and probabilities: SURVEY.md §8(d).

Nothing here is on the timed path; it only builds inputs.
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field
from typing import List, Tuple

import torch

Spot = Tuple[int, int, int]


def shaking_len(n: int) -> int:
    return n * (n + 1) // 2


def shaking_index(i: int, j: int, n: int) -> int:
    return i * n - i * (i - 1) // 2 + (j - i)


@dataclass
class SynthDoc:
    seq_len: int  # N = pair dimension (tokens after the CLS strip)
    n_real: int  # tokens that carry text
    text: List[str]
    bbox: List[List[int]]
    lines: List[Tuple[int, int]]  # inclusive (start, end) of every text line
    entities: List[dict]  # {"lines": [line ids], "label": str}
    spots: List[List[Spot]] = field(default_factory=list)  # 5 lists, return order LE, ELh, ELt, LGh, LGt
    pairs: List[Tuple[int, int]] = field(default_factory=list)  # (question entity id, answer entity id)

    def tags(self) -> List[torch.Tensor]:
        """Dense [P] int64 tags per head (what the collator ships to the model)."""
        out = []
        for sp in self.spots:
            t = torch.zeros(shaking_len(self.seq_len), dtype=torch.int64)
            for i, j, tag in sp:
                t[shaking_index(i, j, self.seq_len)] = tag
            out.append(t)
        return out


def make_document(seq_len: int, doc_id: int = 0, style: str = "rfund") -> SynthDoc:
    rng = random.Random(1000003 * doc_id + seq_len)
    pad = rng.randint(0, 7)
    n_real = max(1, seq_len - 1 - pad)
    lo, hi = (2, 12) if style == "rfund" else (1, 24)
    lines, pos = [], 0
    while pos < n_real:
        ln = min(rng.randint(lo, hi), n_real - pos)
        lines.append((pos, pos + ln - 1))
        pos += ln
    # entities: runs of 1-3 consecutive lines
    entities, li = [], 0
    while li < len(lines):
        r = rng.random()
        k = 1 if r < 0.7 else (2 if r < 0.9 else 3)
        k = min(k, len(lines) - li)
        r = rng.random()
        label = "question" if r < 0.4 else ("answer" if r < 0.8 else ("other" if r < 0.95 else "header"))
        entities.append({"lines": list(range(li, li + k)), "label": label})
        li += k
    le, elh, elt, lgh, lgt = [], [], [], [], []
    for ent in entities:
        if ent["label"] in ("question", "answer"):
            for l in ent["lines"]:
                le.append((lines[l][0], lines[l][1], 1))
            for a, b in zip(ent["lines"][:-1], ent["lines"][1:]):
                lgh.append((lines[a][0], lines[b][0], 1))
                lgt.append((lines[a][1], lines[b][1], 1))
    pairs = []
    for e, ent in enumerate(entities):
        if ent["label"] != "question":
            continue
        nxt = next((f for f in range(e + 1, len(entities)) if entities[f]["label"] == "answer"), None)
        if nxt is None or rng.random() >= 0.8:
            continue
        if any(p[1] == nxt for p in pairs):
            continue  # one question per answer keeps the planted document unambiguous
        pairs.append((e, nxt))
        kh, kt = lines[ent["lines"][0]][0], lines[ent["lines"][-1]][1]
        vh, vt = lines[entities[nxt]["lines"][0]][0], lines[entities[nxt]["lines"][-1]][1]
        if rng.random() < 0.15:
            # value precedes key in reading order -> stored flipped with tag 2.  With a
            # forward-only generator we emulate it by declaring the *answer* the key.
            elh.append((kh, vh, 2))
            elt.append((kt, vt, 2))
        else:
            elh.append((kh, vh, 1))
            elt.append((kt, vt, 1))
    alphabet = "abcdefghijklmnopqrstuvwxyz"
    text = []
    for t in range(seq_len):
        if t < n_real:
            w = "".join(rng.choice(alphabet) for _ in range(rng.randint(1, 5)))
            text.append(w + (" " if rng.random() < 0.5 else ""))
        else:
            text.append("")
    bbox = []
    for t in range(seq_len):
        x0, y0 = rng.randint(0, 900), rng.randint(0, 950)
        bbox.append([x0, y0, x0 + rng.randint(5, 90), y0 + rng.randint(5, 40)])
    return SynthDoc(seq_len, n_real, text, bbox, lines, entities, [le, elh, elt, lgh, lgt], pairs)


def planted_logits(doc: SynthDoc, seed: int = 0, noise: float = 0.5, dtype=torch.float32) -> List[torch.Tensor]:
    """Logits whose decode reproduces the planted document (regime (i)):
    ``noise * N(0,1)``, +4 on class 0 everywhere, +8 on the planted class at planted cells."""
    g = torch.Generator().manual_seed(7919 * seed + doc.seq_len)
    outs = []
    for k, sp in enumerate(doc.spots):
        c = 2 if k == 0 else 3
        z = noise * torch.randn(shaking_len(doc.seq_len), c, generator=g)
        z[:, 0] += 4.0
        for i, j, tag in sp:
            z[shaking_index(i, j, doc.seq_len), tag] += 8.0
        outs.append(z.to(dtype))
    return outs


def hidden_states(batch: int, seq_len: int, hidden: int = 768, doc_id0: int = 0) -> torch.Tensor:
    """Backbone-output stand-in: unit-variance rows, seed 1234 + doc id."""
    xs = []
    for b in range(batch):
        g = torch.Generator().manual_seed(1234 + doc_id0 + b)
        xs.append(torch.randn(seq_len, hidden, generator=g))
    return torch.stack(xs)


def init_decoder_state(
    hin: int = 768,
    hidden: int = 768,
    shrink: bool = True,
    num_layers: int = 2,
    seed: int = 0,
    std: float = 0.02,
    trained_like: bool = False,
    category_weights=(1.0, 10.0, 10.0),
) -> dict:
    """Reference-keyed decoder state dict with the HF-style init PEneoModel applies
    (N(0, 0.02) weights, zero biases; model/modeling_peneo.py:105-106).  ``trained_like``
    scales the output layers x50 and sets b_out[0] = +4 so logits are O(1)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(name, out_f, in_f):
        sd[f"{name}.weight"] = torch.randn(out_f, in_f, generator=g) * std
        sd[f"{name}.bias"] = torch.zeros(out_f)

    d = hidden // 2 if shrink else hin
    if shrink:
        lin("shrink_projection.0", hidden, hin)
        lin("shrink_projection.3", d, hidden)
    lin("handshaking_kernel.combine_fc", d, 2 * d)
    names = ("line_extraction", "ent_linking_h2h", "ent_linking_t2t", "line_grouping_h2h", "line_grouping_t2t")
    for name, c in zip(names, (2, 3, 3, 3, 3)):
        if num_layers == 1:
            lin(f"{name}_fc", c, d)
            last = f"{name}_fc"
        else:
            for l in range(num_layers - 1):
                lin(f"{name}_fc.{3 * l}", d, d)
            last = f"{name}_fc.{3 * (num_layers - 1)}"
            lin(last, c, d)
        if trained_like:
            sd[f"{last}.weight"] *= 50.0
            sd[f"{last}.bias"][0] = 4.0
    w = torch.tensor(list(category_weights), dtype=torch.float32)
    sd["link_loss.weight"] = w.clone()
    sd["le_loss.weight"] = w[:-1].clone()
    return sd


HEAD_NAMES = ("line_extraction", "ent_linking_h2h", "ent_linking_t2t", "line_grouping_h2h", "line_grouping_t2t")


def calibrate_class0_bias(sd: dict, logits: List[torch.Tensor], n: int, spots_per_head: float = 2.0) -> dict:
    """Decode regime (ii) of SURVEY.md §8d ("calibrated random-init"): random-init logits are ~0, so argmax is
    ~uniform and 1/2 - 2/3 of all pairs come out positive — a regime no trained model produces.  Shift the class-0
    output bias of every head to the ``1 - spots_per_head / n`` quantile of ``max_{c>0} l_c - l_0`` of a probe
    document's logits, so that about ``spots_per_head * (n + 1) / 2`` spots per head survive.  ``logits``: the five
    ``[1, P, C]`` tensors the *uncalibrated* weights give for the probe document (any device).  Returns ``sd``
    (modified in place; shipped configuration: the output layer is ``<head>_fc.3``)."""
    q = 1.0 - spots_per_head / n
    for name, lg in zip(HEAD_NAMES, logits):
        lg = lg[0].float()
        margin = lg[:, 1:].max(dim=1)[0] - lg[:, 0]
        shift = torch.quantile(margin[:: max(1, margin.numel() // 100000)], q).item()
        sd[f"{name}_fc.3.bias"][0] += shift
    return sd

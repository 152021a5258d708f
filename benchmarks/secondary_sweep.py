"""SURVEY.md §8(d) secondary sweeps: the configurations next to the shipped one (hidden 768 -> d 384, L = 2), each
through the same plugin call (PEneoDecoderB200.forward in inference mode: the five logits tensors, device-resident inputs):

    Hin = 960 (LiLT width), shrink on, L = 2     -> fused tcgen05 path (K1 takes any Hin)
    L = 1 / L = 3, shrink off (D = Hin = 768)    -> unfused tensor-core forward (pair_heads_generic.cu) and, for
                                                    comparison, the exact fp32 CUDA-core path

One JSON line per configuration.

    python benchmarks/secondary_sweep.py [--seq-len 512] [--batch 8]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from peneo_b200 import PEneoDecoderB200, synth  # noqa: E402


def make_cfg(shrink, layers, precision):
    class Cfg:
        backbone_config = {"hidden_size": 768, "hidden_dropout_prob": 0.1}  # LiLT: input 960 = 768 text + 192 layout
        peneo_decoder_shrink = shrink
        peneo_classifier_num_layers = layers
        peneo_loss_ratio = [1.0] * 5
        peneo_category_weights = [1.0, 10.0, 10.0]
        peneo_ohem_num_positive = -1
        peneo_ohem_num_negative = -1
        inference_mode = True
        peneo_b200_precision = precision

    return Cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seq-len", type=int, default=512)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    n = args.seq_len - 1
    pairs = n * (n + 1) // 2
    cases = [
        ("shipped: Hin 768, shrink, L=2", 768, True, 2, "bf16"),
        ("Hin 960 (LiLT), shrink, L=2", 960, True, 2, "bf16"),
        ("L=1, unfused tensor-core forward", 768, True, 1, "bf16"),
        ("L=1, fp32 CUDA-core path", 768, True, 1, "fp32"),
        ("shrink off (D = 768), L=2, unfused tensor-core forward", 768, False, 2, "bf16"),
        ("shrink off (D = 768), L=2, fp32 CUDA-core path", 768, False, 2, "fp32"),
        ("L=3, unfused tensor-core forward", 768, True, 3, "bf16"),
        ("shipped configuration on the fp32 CUDA-core path", 768, True, 2, "fp32"),
    ]
    for name, hin, shrink, layers, prec in cases:
        d = 384 if shrink else hin
        sd = synth.init_decoder_state(hin=hin, hidden=768, shrink=shrink, num_layers=layers, seed=0, trained_like=True)
        dec = PEneoDecoderB200(make_cfg(shrink, layers, prec), hin)
        dec.load_state_dict(sd)
        dec = dec.to(dev).eval()
        dt = torch.bfloat16 if prec == "bf16" else torch.float32
        xs = [synth.hidden_states(args.batch, n, hin, doc_id0=100 * r).to(dev, dt) for r in range(2)]
        # heads only (token projections + pair heads through PEneoDecoderB200.forward): with uncalibrated weights these
        # configurations emit dense "spots", and the decode of dense garbage would dominate a heads + decode number
        def run(k):
            with torch.no_grad():
                for s in range(k):
                    dec(xs[s % 2])

        run(3)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        run(args.steps)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / args.steps
        per_token = 2.0 * n * ((hin * 768 + 768 * d) if shrink else 0) + 2.0 * n * 2 * d * d
        f_heads = per_token + 10.0 * pairs * d * d * (layers - 1) + 28.0 * pairs * d
        print(json.dumps({"config": name, "hin": hin, "d": d, "layers": layers, "precision": prec, "seq_len": args.seq_len,
                          "batch": args.batch, "ms_per_step": round(ms, 3), "docs_per_s": round(args.batch / (ms * 1e-3), 1),
                          "tflops": round(args.batch * f_heads / (ms * 1e-3) / 1e12, 1)}), flush=True)
        del xs, dec
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

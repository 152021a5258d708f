"""BASELINE configs[3]: decoder-only microbench — pair heads + decode at N = 512 / 1024 / 2048 (N_eff = N - 1),
hidden 768, batch sweep, bf16.  For every (N, batch): device-resident docs/s through HeadsDecodePipeline, the
K2 time (CUDA events) and its fraction of the measured tensor roofline.  One JSON line per point.

    python benchmarks/decoder_microbench.py [--seq-lens 512 1024 2048] [--batches 1 4 16 64 256]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from peneo_b200 import HeadsDecodePipeline, PEneoDecoderB200, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seq-lens", type=int, nargs="+", default=[512, 1024, 2048])
    ap.add_argument("--batches", type=int, nargs="+", default=[1, 4, 16, 64, 256])
    ap.add_argument("--max-logit-gb", type=float, default=24.0)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peak = 1427.8
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:
        pass
    for seq in args.seq_lens:
        n = seq - 1
        pairs = n * (n + 1) // 2
        sd = bench.calibrate_bias(synth.init_decoder_state(seed=0), n, dev)
        dec = PEneoDecoderB200(bench.Cfg, 768)
        dec.load_state_dict(sd)
        dec = dec.to(dev).eval()
        for b in args.batches:
            if b * pairs * 56 / 2**30 > args.max_logit_gb:  # fp32 logits of the batch
                continue
            xs = [synth.hidden_states(b, n, 768, doc_id0=100 * r).to(dev, torch.bfloat16) for r in range(2)]
            texts = [[f"w{t} " for t in range(n)] for _ in range(b)]
            pipe = HeadsDecodePipeline(dec, dev)
            steps = max(8, min(50, int(2e7 / (b * pairs)) + 1))

            def run(k):
                for s in range(k):
                    pipe.submit(xs[s % 2], texts)
                    if len(pipe) >= 2:
                        pipe.result(assemble=False)
                while len(pipe):
                    pipe.result(assemble=False)

            run(10)
            torch.cuda.synchronize()
            pipe.k2_events, pipe.kernel_ms = [], {"k2": []}
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(pipe.compute)
            run(steps)
            t1.record(pipe.compute)
            torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / steps
            k2 = bench.k2_mean_ms(pipe)
            tf = b * (10.0 * pairs * 384 * 384 + 28.0 * pairs * 384) / (k2 * 1e-3) / 1e12
            print(json.dumps({"seq_len": seq, "pair_dim": n, "batch": b, "steps": steps, "ms_per_step": round(ms, 4),
                              "docs_per_s": round(b / (ms * 1e-3), 1), "k2_ms": round(k2, 4), "k2_tflops": round(tf, 1),
                              "k2_frac_of_sustained_peak": round(tf / peak, 4)}), flush=True)
            del pipe, xs
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

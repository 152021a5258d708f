"""BASELINE configs[4]: document-sharded inference over a collection of synthetic forms with mixed sequence lengths
(256..2048, log-uniform, rounded to x8; N_eff = seq - 1), one process per GPU, no data-path collective.

    python benchmarks/mixed_sweep.py --docs 2000                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 \
        benchmarks/mixed_sweep.py --docs 10000                                      # 8 GPUs

Every rank computes the same LPT assignment (peneo_b200.shard), generates the hidden states of ITS documents on its
GPU (synthetic N(0,1), bf16), runs heads + decode through HeadsDecodePipeline grouped by length, and builds the Python
results.  Timed on the device (CUDA events around each rank's work) and by wall clock; the reported docs/s uses the
slowest rank.  The result gather (all_gather_object, as pipeline/evaluation.py does) is outside the timed region.
"""
import argparse
import json
import math
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from peneo_b200 import HeadsDecodePipeline, PEneoDecoderB200, shard, synth  # noqa: E402


def lengths(count, seed=2026):
    rng = random.Random(seed)
    return [max(8, int(round(math.exp(rng.uniform(math.log(256), math.log(2048))) / 8)) * 8) - 1 for _ in range(count)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=2000)
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lens = lengths(args.docs)
    parts = shard.assign_documents(lens, world)
    mine = parts[rank]
    groups = shard.batches_by_length(lens, mine, max_pairs=1 << 23, max_batch=64)
    # one model for every length: class-0 bias calibrated so that about N spots per head survive at the LONGEST length
    # (a fixed per-pair rate, i.e. sparser for short documents — closer to a trained model than a per-length rate)
    sd = bench.calibrate_bias(synth.init_decoder_state(seed=0), 2047, dev)
    dec = PEneoDecoderB200(bench.Cfg, 768)
    dec.load_state_dict(sd)
    dec = dec.to(dev).eval()
    pipe = HeadsDecodePipeline(dec, dev)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)

    def feed():
        for g in groups:
            n = lens[g[0]]
            x = torch.randn(len(g), n, 768, device=dev, generator=gen).to(torch.bfloat16)
            yield x, [[f"w{t} " for t in range(n)]] * len(g)

    # warm-up on a few small groups, then the timed pass over everything
    for i, _ in enumerate(pipe.run((b for j, b in enumerate(feed()) if j < 3), depth=3)):
        pass
    pipe.freeze_host_gc()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(pipe.compute)
    ndocs = npairs = 0
    for g, res in zip(groups, pipe.run(feed(), depth=3)):
        ndocs += len(res)
        npairs += len(g) * shard.pair_cost(lens[g[0]])
    e1.record(pipe.compute)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    stats = torch.tensor([ms, wall * 1e3, float(ndocs), float(npairs)], device=dev, dtype=torch.float64)
    if world > 1:
        allst = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(allst, stats)
    else:
        allst = [stats]
    if rank == 0:
        rows = [s.tolist() for s in allst]
        slow = max(max(r[0], r[1]) for r in rows)
        total_docs, total_pairs = sum(r[2] for r in rows), sum(r[3] for r in rows)
        loads = [shard.rank_load(lens, p) for p in parts]
        print(json.dumps({"what": "mixed-length document-sharded heads+decode (BASELINE configs[4])", "n_gpus": world,
                          "docs": int(total_docs), "pairs": int(total_pairs), "ms_slowest_rank": slow,
                          "docs_per_s": total_docs / (slow * 1e-3), "gpairs_per_s": total_pairs / (slow * 1e-3) / 1e9,
                          "rank_ms_device": [round(r[0], 1) for r in rows], "rank_ms_wall": [round(r[1], 1) for r in rows],
                          "load_imbalance_max_over_mean": max(loads) / (sum(loads) / world)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

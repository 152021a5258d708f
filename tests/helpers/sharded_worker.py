"""torchrun worker: document-sharded heads+decode on N GPUs must reproduce the single-GPU results.
Launched by tests/test_gpu_multi.py (one process per GPU, NCCL)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from peneo_b200 import PEneoDecoderB200, shard, synth  # noqa: E402


class Cfg:
    backbone_config = {"hidden_size": 768, "hidden_dropout_prob": 0.1}
    peneo_decoder_shrink = True
    peneo_classifier_num_layers = 2
    peneo_loss_ratio = [1.0] * 5
    peneo_category_weights = [1.0, 10.0, 10.0]
    peneo_ohem_num_positive = -1
    peneo_ohem_num_negative = -1
    inference_mode = False
    peneo_b200_precision = "bf16"


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    sd = synth.init_decoder_state(seed=3, trained_like=True)
    dec = PEneoDecoderB200(Cfg, 768)
    dec.load_state_dict(sd)
    dec = dec.to(dev).eval()

    # ---- inference: mixed lengths, every rank holds the same collection, decodes only its share
    lengths = [63, 127, 63, 255, 95, 127, 63, 191, 95, 255, 63, 127]
    hidden = [synth.hidden_states(1, n, 768, doc_id0=900 + i)[0].to(torch.bfloat16) for i, n in enumerate(lengths)]
    texts = [[f"d{i}t{t} " for t in range(n)] for i, n in enumerate(lengths)]
    mine = shard.sharded_decode(dec, hidden, texts)
    merged = shard.gather_results(mine)
    assert sorted(merged) == list(range(len(lengths)))
    if rank == 0:
        single = shard.sharded_decode(dec, hidden, texts, rank=0, world_size=1)
        for i in range(len(lengths)):
            assert merged[i][0] == single[i][0] and merged[i][1] == single[i][1], i
            for a, b in zip(merged[i][2:], single[i][2:]):
                assert list(a.items()) == list(b.items()), i

    # ---- training: per-rank batch, NCCL gradient average == mean of the per-rank gradients
    n, b = 31, 2
    x = synth.hidden_states(b, n, 768, doc_id0=100 * rank).to(dev)
    docs = [synth.make_document(n, doc_id=700 + 10 * rank + i) for i in range(b)]
    tags = [torch.stack([d.tags()[k] for d in docs]).to(dev) for k in range(5)]
    dec.zero_grad(set_to_none=True)
    dec(x, None, *tags).loss.backward()
    local_grads = {k: p.grad.clone() for k, p in dec.named_parameters()}
    nbytes = shard.allreduce_gradients(dec.parameters())
    assert nbytes == sum(p.numel() for p in dec.parameters()) * 4
    gathered = [None] * world
    dist.all_gather_object(gathered, {k: v.cpu() for k, v in local_grads.items()})
    for k, p in dec.named_parameters():
        mean = sum(g[k] for g in gathered) / world
        err = (p.grad.cpu() - mean).abs().max().item() / max(mean.abs().max().item(), 1e-30)
        assert err <= 1e-5, (k, err)
    # ---- evaluation metrics ("next" row N2): the fixed-width int64 all_gather under NCCL gives every rank the
    #      numbers the reference computes in one process over all files (pipeline/evaluation.py:150-183, 416-513),
    #      including the de-duplication of files the distributed sampler put on two ranks
    from peneo_b200 import evaluation as ev

    g = torch.load(os.path.join(ROOT, "tests", "golden", "eval.pt"), weights_only=False)
    nfiles = len(g["names"])
    idx = [i for i in range(nfiles) if i % world == rank] + ([0] if rank == world - 1 else [])  # file 0 twice
    sub = lambda xs: [xs[i] for i in idx]  # noqa: E731
    kv = ev.calculate_KVPE_metric(sub(g["preds"]), sub(g["gts"]), sub(g["names"]))
    det = ev.calculate_detail_KVPE_metric(sub(g["preds"]), sub(g["gts"]), sub(g["names"]))
    assert kv[0] == g["kvpe"][0], (kv[0], g["kvpe"][0])
    assert det[0] == g["detail"][0]
    assert kv[1]["num_sample_processed"] == g["kvpe"][1]["num_sample_processed"]

    # ---- DDP (what HF Trainer wraps the reference's model in): same averaged gradients as the flat bucket above
    ddp = torch.nn.parallel.DistributedDataParallel(dec, device_ids=[local])
    dec.zero_grad(set_to_none=True)
    ddp(x, None, *tags).loss.backward()
    for k, p in dec.named_parameters():
        mean = sum(gr[k] for gr in gathered) / world
        err = (p.grad.cpu() - mean).abs().max().item() / max(mean.abs().max().item(), 1e-30)
        # (a second backward pass: the bf16 kernels accumulate with fp32 atomics, so two runs differ in the last bits)
        assert err <= 1e-3, ("ddp", k, err)
    dist.barrier()
    if rank == 0:
        print(f"sharded ok world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

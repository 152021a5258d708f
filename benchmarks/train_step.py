"""Fine-tuning step of the decoder (BASELINE configs[2] shape): forward + fused loss + backward on synthetic
SIBR-shaped documents, timed with CUDA events.  Reports ms / step and the fraction of the tensor roofline
against 3 x F_heads (what autograd would execute with stored activations).

    python benchmarks/train_step.py --seq-len 1024 --batch 4 --precision bf16
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 benchmarks/train_step.py     # DDP: NCCL gradient all-reduce
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from peneo_b200 import PEneoDecoderB200, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seq-len", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=12)  # > 10: DDP synchronises host and GPU in its first 10 iterations
    ap.add_argument("--hin", type=int, default=768, help="width of the backbone output (960 = LiLT, BASELINE configs[2])")
    ap.add_argument("--no-fused-loss", action="store_true", help="separate loss kernels + explicit dlogits (A/B)")
    ap.add_argument("--no-shrink", action="store_true", help="peneo_decoder_shrink = False (D = hin): unfused tensor-core route")
    ap.add_argument("--layers", type=int, default=2, help="peneo_classifier_num_layers")
    args = ap.parse_args()

    class Cfg:
        backbone_config = {"hidden_size": 768, "hidden_dropout_prob": 0.1}
        peneo_decoder_shrink = not args.no_shrink
        peneo_classifier_num_layers = args.layers
        peneo_loss_ratio = [1.0] * 5
        peneo_category_weights = [1.0, 10.0, 10.0]
        peneo_ohem_num_positive = -1
        peneo_ohem_num_negative = -1
        inference_mode = False
        peneo_b200_precision = args.precision

    n = args.seq_len - 1
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dec = PEneoDecoderB200(Cfg, args.hin)
    dec.load_state_dict(synth.init_decoder_state(args.hin, 768, not args.no_shrink, args.layers, seed=0))
    dec = dec.cuda().eval()
    dec.fused_loss = not args.no_fused_loss
    module = dec
    if world > 1:  # what HF Trainer does for the reference: DDP, one process per GPU, gradients averaged over NCCL
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # broadcast_buffers=False: only constant buffers (class weights); the default makes DDP.forward a host-GPU sync
        dec = torch.nn.parallel.DistributedDataParallel(dec, device_ids=[local_rank], broadcast_buffers=False)
    x = synth.hidden_states(args.batch, n, args.hin, doc_id0=1000 * rank).cuda().requires_grad_(True)
    docs = [synth.make_document(n, doc_id=1000 * rank + i, style="sibr") for i in range(args.batch)]
    tags = [torch.stack([d.tags()[k] for d in docs]).cuda() for k in range(5)]

    def step():
        module.zero_grad(set_to_none=True)
        x.grad = None
        out = dec(x, None, *tags)
        out.loss.backward()
        return out.loss

    for _ in range(args.warmup):
        step()
    import gc

    gc.collect()
    gc.freeze()  # a full CPython collection (~45 ms with torch loaded) inside a step stalls every rank of a DDP job
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for _ in range(args.steps):
        loss = step()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / args.steps
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    p = n * (n + 1) // 2
    d = args.hin if args.no_shrink else 384
    f_tok = 2.0 * n * ((0 if args.no_shrink else args.hin * 768 + 768 * 384) + 2 * d * d)
    f_heads = f_tok + 10.0 * (args.layers - 1) * p * d * d + 28.0 * p * d
    peak = 1427.8
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:
        pass
    tf = 3.0 * f_heads * args.batch / (ms * 1e-3) / 1e12  # per GPU
    if rank == 0:
        print(json.dumps({"what": "decoder fine-tuning step (fwd + loss + bwd" + (", DDP all-reduce)" if world > 1 else ")"),
                      "precision": args.precision, "fused_loss": not args.no_fused_loss, "hin": args.hin, "shrink": not args.no_shrink, "layers": args.layers, "seq_len": args.seq_len, "n_gpus": world,
                      "batch": args.batch, "ms_per_step": ms, "docs_per_s": world * args.batch / (ms * 1e-3), "loss": float(loss.detach()),
                      "tflops_vs_3F": tf, "frac_of_sustained_bf16_peak": tf / peak,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

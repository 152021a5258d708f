"""torchrun --nproc-per-node 2 benchmarks/scratch/ddp_variants.py : where does the DDP step lose time?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from peneo_b200 import PEneoDecoderB200, synth

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

world, rank, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n, batch = 511, 32
x = synth.hidden_states(batch, n, 768, doc_id0=1000 * rank).cuda().requires_grad_(True)
docs = [synth.make_document(n, doc_id=1000 * rank + i, style="rfund") for i in range(batch)]
tags = [torch.stack([d.tags()[k] for d in docs]).cuda() for k in range(5)]

def run(name, wrap, post=None, steps=12, warm=4):
    dec = PEneoDecoderB200(Cfg, 768)
    dec.load_state_dict(synth.init_decoder_state(768, 768, True, 2, seed=0))
    dec = dec.cuda().eval()
    m = wrap(dec)
    def step():
        dec.zero_grad(set_to_none=True)
        x.grad = None
        o = m(x, None, *tags)
        o.loss.backward()
        if post: post(dec)
    for _ in range(warm): step()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(steps): step()
    e1.record(); t_host = (time.perf_counter() - t0) / steps * 1e3
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if rank == 0: print(f"{name:40s} {ms:8.2f} ms/step   host enqueue {t_host:6.2f} ms", flush=True)
    del m, dec
    torch.cuda.empty_cache()

DDP = torch.nn.parallel.DistributedDataParallel
run("no wrapper, no all-reduce", lambda d: d)
if world > 1:
    def flat(dec):
        gs = [p.grad for p in dec.parameters() if p.grad is not None]
        buf = torch.cat([g.reshape(-1) for g in gs]); dist.all_reduce(buf); buf /= world
    run("manual flat all-reduce after backward", lambda d: d, flat)
    run("DDP default", lambda d: DDP(d, device_ids=[lr]))
    run("DDP broadcast_buffers=False", lambda d: DDP(d, device_ids=[lr], broadcast_buffers=False))
    run("DDP bucket_view, no buffers", lambda d: DDP(d, device_ids=[lr], broadcast_buffers=False, gradient_as_bucket_view=True))
    run("DDP static_graph", lambda d: DDP(d, device_ids=[lr], broadcast_buffers=False, static_graph=True))
    dist.destroy_process_group()

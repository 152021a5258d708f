"""torchrun --nproc-per-node 2 benchmarks/scratch/ddp_timeline.py : per-step GPU / host timeline of the DDP training step"""
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
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n, batch = 511, 32
x = synth.hidden_states(batch, n, 768, doc_id0=1000 * rank).cuda().requires_grad_(True)
docs = [synth.make_document(n, doc_id=1000 * rank + i, style="sibr") for i in range(batch)]
tags = [torch.stack([d.tags()[k] for d in docs]).cuda() for k in range(5)]
dec = PEneoDecoderB200(Cfg, 768)
dec.load_state_dict(synth.init_decoder_state(768, 768, True, 2, seed=0))
dec = dec.cuda().eval()
m = torch.nn.parallel.DistributedDataParallel(dec, device_ids=[lr])
steps = 14
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
host = []
for s in range(steps):
    t0 = time.perf_counter()
    dec.zero_grad(set_to_none=True); x.grad = None
    ev[s][0].record()
    o = m(x, None, *tags)
    ev[s][1].record()
    t1 = time.perf_counter()
    o.loss.backward()
    ev[s][2].record()
    t2 = time.perf_counter()
    host.append((t0, t1, t2))
torch.cuda.synchronize()
for r in range(world):
    dist.barrier()
    if r == rank:
        for s in range(4, steps):
            nxt = ev[s + 1][0] if s + 1 < steps else None
            print(f"rank {rank} step {s}: gpu fwd {ev[s][0].elapsed_time(ev[s][1]):6.2f} bwd(+allreduce) {ev[s][1].elapsed_time(ev[s][2]):6.2f}"
                  f" gap->next {ev[s][2].elapsed_time(nxt) if nxt else 0:6.2f} | host fwd {1e3*(host[s][1]-host[s][0]):6.2f} bwd {1e3*(host[s][2]-host[s][1]):6.2f}", flush=True)
dist.destroy_process_group()

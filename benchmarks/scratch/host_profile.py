"""Host-side cost of the training step: python benchmarks/scratch/host_profile.py"""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
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

n, batch = 511, 32
x = synth.hidden_states(batch, n, 768).cuda().requires_grad_(True)
docs = [synth.make_document(n, doc_id=i, style="rfund") for i in range(batch)]
tags = [torch.stack([d.tags()[k] for d in docs]).cuda() for k in range(5)]
dec = PEneoDecoderB200(Cfg, 768)
dec.load_state_dict(synth.init_decoder_state(768, 768, True, 2, seed=0))
dec = dec.cuda().eval()
def step():
    dec.zero_grad(set_to_none=True)
    x.grad = None
    o = dec(x, None, *tags)
    o.loss.backward()
for _ in range(4): step()
torch.cuda.synchronize()
# phases
for name in ("fwd", "bwd"):
    pass
t = []
for _ in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dec.zero_grad(set_to_none=True); x.grad = None
    o = dec(x, None, *tags)
    t1 = time.perf_counter()
    o.loss.backward()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    t.append((t1 - t0, t2 - t1, t3 - t2))
print("host ms: fwd %.2f bwd %.2f  then wait for GPU %.2f" % tuple(1e3 * sum(c) / len(t) for c in zip(*t)))
pr = cProfile.Profile()
torch.cuda.synchronize()
pr.enable()
for _ in range(5): step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)

"""Where the MMA-issuing thread of K2 waits (needs a library built with PENEO_NVCC_EXTRA=-DPENEO_K2_PROFILE):
    PENEO_NVCC_EXTRA=-DPENEO_K2_PROFILE python -m peneo_b200.build --force && python benchmarks/scratch/k2_profile.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from peneo_b200 import _lib, synth

lib = _lib.load()
out = (C.c_ulonglong * 8)()
dev = torch.device("cuda", 0)
n = 511
dec, _ = bench.make_decoder(n, dev)
xs = [synth.hidden_states(32, n, 768, doc_id0=100 * r).to(device=dev, dtype=torch.bfloat16) for r in range(4)]
with torch.no_grad():
    for r in range(4): dec(xs[r])
    torch.cuda.synchronize()
    lib.peneo_debug_k2_profile(out, 1)
    for r in range(8): dec(xs[r % 4])
    torch.cuda.synchronize()
lib.peneo_debug_k2_profile(out, 0)
v = list(out)
ctas, tot = v[6], v[5]
names = ["S chunk full", "W stage full", "m ready", "z free", "W_out stage full"]
print(f"leader-CTA launches {ctas}, issuing loop {tot / ctas / 1e6:.3f} Mcycles per CTA and launch")
for nme, c in zip(names, v[:5]): print(f"  waiting for {nme:18s} {100.0 * c / tot:6.2f} % of the issuing thread's time")
print(f"  issuing / not waiting         {100.0 * (tot - sum(v[:5])) / tot:6.2f} %")

import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from peneo_b200 import HeadsDecodePipeline, synth
dev = torch.device("cuda", 0)
for seq, b in ((2048, 2), (1024, 8), (512, 32)):
    n = seq - 1
    dec, _ = bench.make_decoder(n, dev)
    xs = [synth.hidden_states(b, n, 768, doc_id0=100 * r).to(dev, torch.bfloat16) for r in range(2)]
    texts = [[f"w{t} " for t in range(n)] for _ in range(b)]
    pipe = HeadsDecodePipeline(dec, dev)
    def run(k, rec=None):
        for s in range(k):
            t0 = time.perf_counter(); pipe.submit(xs[s % 2], texts); t1 = time.perf_counter()
            if len(pipe) >= 3:
                pipe.result(assemble=False)
            t2 = time.perf_counter()
            if rec is not None: rec.append((t1 - t0, t2 - t1))
        while len(pipe): pipe.result(assemble=False)
    run(3); torch.cuda.synchronize()
    rec = []
    t0 = time.perf_counter(); run(10, rec); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(seq, b, "wall ms/step", (t1 - t0) * 100, "submit ms", [round(a * 1e3, 2) for a, _ in rec], "result ms", [round(c * 1e3, 2) for _, c in rec])
    del pipe

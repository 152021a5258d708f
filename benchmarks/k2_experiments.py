"""Bottleneck study of K2 (pair_heads_tc_kernel): times the kernel under the PENEO_K2_DBG experiment bits
(1 = W_mid streaming disabled after the first tile, 2 = epilogue without the SiLU math; results are wrong when
a bit is set, only the timing is meaningful).  GPU box only:  python benchmarks/k2_experiments.py 0 1 2 3"""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from peneo_b200 import PEneoDecoderB200, ops, synth
    import bench
    n, b = int(sys.argv[2]), int(sys.argv[3])
    sd = synth.init_decoder_state(seed=0)
    dec = PEneoDecoderB200(bench.Cfg, 768); dec.load_state_dict(sd); dec = dec.cuda().eval()
    x = synth.hidden_states(b, n, 768).cuda().bfloat16()
    pack = dec._weight_pack(x.device)
    ab = ops.token_projections(pack, x)
    for _ in range(3): ops.pair_heads(pack, ab, b, n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(int(os.environ.get("K2_ITERS", "10"))): ops.pair_heads(pack, ab, b, n)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / int(os.environ.get("K2_ITERS", "10"))
    p = n * (n + 1) // 2
    fl = b * (10.0 * p * 384 * 384 + 28.0 * p * 384)
    print(json.dumps({"dbg": os.environ.get("PENEO_K2_DBG", "0"), "n": n, "b": b, "ms": ms, "tflops": fl / ms / 1e9}))
else:
    for dbg in sys.argv[1:] or ["0", "1", "2", "3"]:
        env = dict(os.environ, PENEO_K2_DBG=dbg)
        subprocess.run([sys.executable, __file__, "child", "511", "32"], env=env)

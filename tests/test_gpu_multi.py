"""Multi-GPU check (needs >= 2 GPUs: run with `gpurun --gpus N -- python -m pytest tests -m gpu`): one process per
VISIBLE GPU under torchrun, NCCL; document-sharded inference equals the single-GPU result, the gradient all-reduce
(flat bucket and DDP) equals the mean of the per-rank gradients, and the evaluation-metric all_gather gives every
rank the single-process numbers."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_sharded_inference_and_gradient_allreduce_nccl():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = n  # every visible GPU
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import socket

    with socket.socket() as sock:  # a free rendezvous port
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(root, "tests", "helpers", "sharded_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert f"sharded ok world={world}" in out.stdout

"""Host-side multi-GPU logic on CPU: deterministic document sharding, result gather and the gradient
all-reduce, exercised with world_size-2 gloo process groups (no GPU)."""
import os
import random
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from peneo_b200 import shard


def _mixed_lengths(count=10000, seed=2026):
    """BASELINE configs[4]: mixed sequence lengths 256..2048, log-uniform, rounded to x8, minus CLS."""
    import math

    rng = random.Random(seed)
    out = []
    for _ in range(count):
        s = math.exp(rng.uniform(math.log(256), math.log(2048)))
        out.append(max(8, int(round(s / 8)) * 8) - 1)
    return out


def test_lpt_assignment_is_a_balanced_partition():
    lengths = _mixed_lengths()
    for world in (1, 2, 4, 8):
        parts = shard.assign_documents(lengths, world)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(len(lengths)))  # every document exactly once
        loads = [shard.rank_load(lengths, p) for p in parts]
        assert max(loads) - min(loads) <= shard.pair_cost(max(lengths))  # LPT bound: within one largest job
        assert max(loads) / (sum(loads) / world) < 1.001
    assert shard.assign_documents([], 4) == [[], [], [], []]
    assert shard.assign_documents([5], 2) == [[0], []]


def test_batches_group_equal_lengths_and_respect_bounds():
    lengths = _mixed_lengths(500)
    docs = shard.assign_documents(lengths, 2)[1]
    groups = shard.batches_by_length(lengths, docs, max_pairs=1 << 21, max_batch=16)
    assert sorted(i for g in groups for i in g) == sorted(docs)
    for g in groups:
        assert len({lengths[i] for i in g}) == 1 and len(g) <= 16
        assert len(g) == 1 or len(g) * shard.pair_cost(lengths[g[0]]) <= 1 << 21


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lengths = _mixed_lengths(300, seed=7)
        mine = shard.assign_documents(lengths, world)[rank]
        # stand-in for the per-document decode result; document 0 is duplicated on rank 1 (sampler padding)
        local = {i: ("doc", i, lengths[i], rank) for i in mine}
        if rank == 1:
            local[0] = ("doc", 0, lengths[0], rank)
        merged = shard.gather_results(local)
        ok_cover = sorted(merged) == list(range(len(lengths)))
        owner0 = merged[0][3]
        # gradient all-reduce: rank r holds grad = (r + 1) * ones -> average = (1 + 2) / 2
        lin = torch.nn.Linear(4, 3)
        for p in lin.parameters():
            p.grad = torch.full_like(p, float(rank + 1))
        nbytes = shard.allreduce_gradients(lin.parameters())
        avg_ok = all(torch.allclose(p.grad, torch.full_like(p, 1.5)) for p in lin.parameters())
        q.put((rank, ok_cover, owner0, avg_ok, nbytes, shard.rank_load(lengths, mine)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_gather_and_gradient_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    rows = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rows.sort()
    for rank, ok_cover, owner0, avg_ok, nbytes, load in rows:
        assert ok_cover and avg_ok and nbytes == (4 * 3 + 3) * 4
    # the duplicated document keeps the copy of the rank that owns it in the assignment order of ranks
    assert rows[0][2] == rows[1][2]
    loads = [r[5] for r in rows]
    assert abs(loads[0] - loads[1]) <= shard.pair_cost(2047)

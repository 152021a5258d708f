"""Document sharding across the GPUs of one box (one process per GPU).

Every document — and every pair inside it — is independent in the heads, the per-document part of the
loss and the decode, so the path shards **by document** with no data-path collective:

* inference: each rank decodes its own documents; results are exchanged once, after the timed work,
  the way the reference's evaluation does (``all_gather_object`` of per-file rows,
  pipeline/evaluation.py:150-156);
* training: the only collective is the gradient all-reduce DDP already performs for the reference
  (README.md:218: ``torchrun --nproc_per_node``).  ``allreduce_gradients`` is that step for callers that
  drive the decoder without DDP: one flat bucket, NCCL over NVLink (gloo on CPU for the tests), averaged
  over ranks — per-rank losses are batch-global weighted means and the *gradients* are averaged, exactly
  as under DDP.

Cost model: the pair kernels do work proportional to ``P = N (N + 1) / 2``.
"""
from __future__ import annotations

import heapq
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def pair_cost(n: int) -> int:
    return n * (n + 1) // 2


def assign_documents(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first: documents sorted by pair count, each given to the least loaded
    rank.  Deterministic (ties -> lower document index, lower rank), so every rank can compute the whole
    assignment locally without communication.  Returns ``world_size`` lists of document indices, each
    sorted by (length, index) so that equal-length documents sit next to each other for batching."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(lengths)), key=lambda i: (-pair_cost(int(lengths[i])), i))
    heap = [(0, r) for r in range(world_size)]
    heapq.heapify(heap)
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        load, r = heapq.heappop(heap)
        out[r].append(i)
        heapq.heappush(heap, (load + pair_cost(int(lengths[i])), r))
    for r in range(world_size):
        out[r].sort(key=lambda i: (int(lengths[i]), i))
    return out


def rank_load(lengths: Sequence[int], docs: Iterable[int]) -> int:
    return sum(pair_cost(int(lengths[i])) for i in docs)


def batches_by_length(lengths: Sequence[int], docs: Sequence[int], max_pairs: int = 1 << 23,
                      max_batch: int = 256) -> List[List[int]]:
    """Group a rank's documents into kernel batches of equal length N (the kernels take [B, N, Hin]),
    each bounded by ``max_pairs`` total pairs (logits are 56 B / pair) and ``max_batch`` documents."""
    out: List[List[int]] = []
    cur: List[int] = []
    cur_n = None
    for i in sorted(docs, key=lambda i: (int(lengths[i]), i)):
        n = int(lengths[i])
        if cur and (n != cur_n or len(cur) >= max_batch or (len(cur) + 1) * pair_cost(n) > max_pairs):
            out.append(cur)
            cur = []
        cur.append(i)
        cur_n = n
    if cur:
        out.append(cur)
    return out


def gather_results(local: Dict[int, object], group=None) -> Dict[int, object]:
    """All ranks' ``{doc index: result}`` maps merged on every rank (untimed epilogue; plain Python objects,
    like pipeline/evaluation.py:150-156).  A document present on several ranks (sampler padding) keeps the
    lowest rank's copy, the reference's de-duplication rule (evaluation.py:173-175)."""
    if not (dist.is_available() and dist.is_initialized()):
        return dict(local)
    world = dist.get_world_size(group)
    parts: List[Optional[dict]] = [None] * world
    dist.all_gather_object(parts, local, group=group)
    merged: Dict[int, object] = {}
    for part in parts:
        for k, v in part.items():
            merged.setdefault(k, v)
    return merged


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None, average: bool = True) -> int:
    """Sum (or average) the ``.grad`` of ``params`` over all ranks through one flat bucket.
    Returns the number of bytes reduced (0 when not running distributed)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not (dist.is_available() and dist.is_initialized()):
        return 0
    world = dist.get_world_size(group)
    # one flat fp32 bucket: a single cat, ONE collective (NCCL averages inside the all-reduce), one fused copy back
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    nccl = dist.get_backend(group) == "nccl"
    dist.all_reduce(flat, op=dist.ReduceOp.AVG if (average and nccl) else dist.ReduceOp.SUM, group=group)
    if average and not nccl:
        flat /= world
    views, off = [], 0
    for g in grads:
        k = g.numel()
        views.append(flat[off : off + k].view_as(g))
        off += k
    if all(g.dtype == torch.float32 for g in grads):
        torch._foreach_copy_(grads, views)
    else:
        for g, v in zip(grads, views):
            g.copy_(v)
    return flat.numel() * 4


def sharded_decode(decoder, hidden: Sequence[torch.Tensor], texts: Sequence[List[str]], rank: Optional[int] = None,
                   world_size: Optional[int] = None, depth: int = 2) -> Dict[int, Tuple]:
    """Heads + decode of this rank's share of a document collection.  ``hidden[i]``: [N_i, Hin] (host or
    device).  Returns ``{doc index: reference 7-tuple}`` for the documents assigned to this rank."""
    from .pipeline import HeadsDecodePipeline

    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    lengths = [int(h.shape[0]) for h in hidden]
    mine = assign_documents(lengths, world_size)[rank]
    pipe = HeadsDecodePipeline(decoder)
    groups = batches_by_length(lengths, mine)

    def feed():
        for g in groups:
            x = torch.stack([hidden[i] for i in g])
            if not x.is_cuda:
                x = x.pin_memory()
            yield x, [texts[i] for i in g]

    out: Dict[int, Tuple] = {}
    for g, res in zip(groups, pipe.run(feed(), depth=depth)):
        for i, r in zip(g, res):
            out[i] = r
    return out

"""pipeline/evaluation.py drop-ins ("next" row N2) against outputs of the real reference
(tests/golden/eval.pt, written by oracle/make_golden.py) and across a world_size-2 gloo group."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from peneo_b200 import evaluation as ev


def test_kvpe_metrics_equal_reference_objects(golden):
    g = golden("eval.pt")
    got = ev.calculate_KVPE_metric(g["preds"], g["gts"], g["names"])
    assert got == g["kvpe"]
    assert list(got[0]) == list(g["kvpe"][0]) and list(got[1]) == list(g["kvpe"][1])  # key order too
    got = ev.calculate_detail_KVPE_metric(g["preds"], g["gts"], g["names"])
    assert got == g["detail"]
    assert list(got[0]) == list(g["detail"][0]) and list(got[1]) == list(g["detail"][1])
    assert all(isinstance(v, float) for v in got[0].values())
    assert 0.0 < got[0]["f1"] < 1.0 and got[1]["kv_pair"]["num_gt"] > 0  # the fixture exercises TP, FP and FN


def test_unhashable_items_fall_back_to_list_membership():
    pred = [("k", "v", [0, 0, 1, 1], [2, 2, 3, 3]), ("a", "b", [0, 0, 1, 1], [9, 9, 9, 9])]
    gt = [("k", "v", [0, 0, 1, 1], [2, 2, 3, 3]), ("c", "d", [1, 1, 2, 2], [2, 2, 3, 3])]
    p, r, f1, n_pred, n_gt, n_ok, detail = ev._calculate_KV_metric_core(pred, gt, return_detail=True)
    assert (n_pred, n_gt, n_ok) == (2.0, 2.0, 1.0) and p == 0.5 and r == 0.5
    assert [d["status"] for d in detail] == ["TP", "FP", "FN"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, golden_path, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.load(golden_path, weights_only=False)
        n = len(g["names"])
        # rank 0: files 0..3 ; rank 1: files 3..n-1 (file 3 on both ranks: sampler padding)
        idx = list(range(0, 4)) if rank == 0 else list(range(3, n))
        sub = lambda xs: [xs[i] for i in idx]  # noqa: E731
        kv = ev.calculate_KVPE_metric(sub(g["preds"]), sub(g["gts"]), sub(g["names"]))
        det = ev.calculate_detail_KVPE_metric(sub(g["preds"]), sub(g["gts"]), sub(g["names"]))
        q.put((rank, kv[0], kv[1]["num_sample_processed"], det[0]))
    finally:
        dist.destroy_process_group()


def test_cross_rank_gather_matches_single_process(golden):
    g = golden("eval.pt")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval.pt")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, path, q)) for r in range(2)]
    for p in procs:
        p.start()
    rows = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # every rank sees the global numbers; they equal the reference's single-process numbers over all files
    # (both de-duplicate by file name)
    for _rank, kv, processed, det in rows:
        assert kv == g["kvpe"][0]
        assert det == g["detail"][0]
        assert processed == g["kvpe"][1]["num_sample_processed"]

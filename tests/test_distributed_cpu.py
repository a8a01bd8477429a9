"""World-size-2 gloo test (CPU) of the multi-GPU host logic: range sharding of the catalogue, all-gather layout of the
per-shard top-k lists and the merge order rule — with the CUDA ops replaced by injected oracle-backed stand-ins (the
product has no CPU path; only the plumbing in poi_recommendation_models_b200/distributed.py runs here)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _merge_ref(gs, gi):
    """[U,L,k] -> [U,k] by (score desc, id asc), -1 ids are padding (the rule of nais_topk_merge)."""
    U, L, k = gs.shape
    s, i = gs.reshape(U, L * k).numpy(), gi.reshape(U, L * k).numpy()
    out_s, out_i = np.full((U, k), -np.inf, np.float32), np.full((U, k), -1, np.int32)
    for u in range(U):
        valid = i[u] >= 0
        order = np.lexsort((i[u][valid], -s[u][valid]))[:k]
        out_s[u, :len(order)], out_i[u, :len(order)] = s[u][valid][order], i[u][valid][order]
    return torch.from_numpy(out_s), torch.from_numpy(out_i)


def _worker(rank, world, port, q):
    try:
        _worker_body(rank, world, port, q)
    except Exception:  # surface the traceback in the parent
        import traceback
        q.put((rank, traceback.format_exc()))


def _worker_body(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from types import SimpleNamespace
    from poi_recommendation_models_b200.distributed import ShardedRanker, allreduce_gradients
    N, U, k = 1000, 6, 20
    rng = np.random.default_rng(0)  # same scores on every rank
    scores = rng.normal(size=(U, N)).astype(np.float32)
    scores[:, ::7] = scores[:, 3:4]  # ties across shards: the id rule must break them identically

    def local_topk(users, kk, lo, hi, precision):
        s = torch.from_numpy(scores[:, lo:hi])
        kk2 = min(kk, hi - lo)
        out_s, out_i = torch.full((U, kk), -float("inf")), torch.full((U, kk), -1, dtype=torch.int32)
        for u in range(U):
            order = np.lexsort((np.arange(lo, hi), -s[u].numpy()))[:kk2]
            out_s[u, :kk2], out_i[u, :kk2] = s[u][order], torch.from_numpy((order + lo).astype(np.int32))
        return out_s, out_i

    model = SimpleNamespace(item_num=N)
    r = ShardedRanker(model, rank, world, local_topk=local_topk, merge=_merge_ref)
    users = SimpleNamespace(offsets=torch.zeros(1))
    s, i = r.topk(users, k)
    ref_order = [np.lexsort((np.arange(N), -scores[u]))[:k] for u in range(U)]
    ok = all(np.array_equal(i[u].numpy(), ref_order[u].astype(np.int32)) for u in range(U))
    ok = ok and all(np.array_equal(s[u].numpy(), scores[u][ref_order[u]]) for u in range(U))
    ok = ok and (r.lo, r.hi) == ((0, 512) if rank == 0 else (512, 1000))
    # data-parallel gradient averaging
    lin = torch.nn.Linear(4, 3)
    for p_ in lin.parameters():
        p_.grad = torch.full_like(p_, float(rank + 1))
    allreduce_gradients(lin, world)
    ok = ok and all(torch.allclose(p_.grad, torch.full_like(p_, 1.5)) for p_ in lin.parameters())
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_ranker_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]

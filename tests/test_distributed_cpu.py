"""World-size-2 / 4 gloo tests (CPU) of the multi-GPU host logic: range sharding of the catalogue, user slicing of the
batch, all-gather layout of the per-rank top-k lists and the merge order rule — with the CUDA ops replaced by injected oracle-backed stand-ins (the
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


def _worker(rank, world, port, q, min_shard, want_grid):
    try:
        _worker_body(rank, world, port, q, min_shard, want_grid)
    except Exception:  # surface the traceback in the parent
        import traceback
        q.put((rank, traceback.format_exc()))


def _worker_body(rank, world, port, q, min_shard, want_grid):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from types import SimpleNamespace
    from poi_recommendation_models_b200 import ops
    from poi_recommendation_models_b200.distributed import ShardedRanker, allreduce_gradients, shard_range
    N, U, k = 1000, 7, 20  # 7 users: the last user slice is ragged and gets padded for the gather
    rng = np.random.default_rng(0)  # same scores on every rank
    scores = rng.normal(size=(U, N)).astype(np.float32)
    scores[:, ::7] = scores[:, 3:4]  # ties across shards: the id rule must break them identically

    def local_topk(users, kk, lo, hi, precision):
        rows = users.ids
        kk2 = min(kk, hi - lo)
        out_s, out_i = torch.full((len(rows), kk), -float("inf")), torch.full((len(rows), kk), -1, dtype=torch.int32)
        for j, u in enumerate(rows):
            s = scores[u, lo:hi]
            order = np.lexsort((np.arange(lo, hi), -s))[:kk2]
            out_s[j, :kk2], out_i[j, :kk2] = torch.from_numpy(s[order]), torch.from_numpy((order + lo).astype(np.int32))
        return out_s, out_i

    def slice_users(users, u0, u1):
        return SimpleNamespace(offsets=users.offsets, n_users=u1 - u0, ids=users.ids[u0:u1])

    model = SimpleNamespace(item_num=N)
    def merge(keys):  # [L, U, k] packed keys, the layout of the single all-gather -> reference merge of the unpacked lists
        gs, gi = ops.keys_to_lists(keys)
        return _merge_ref(gs.permute(1, 0, 2).contiguous(), gi.permute(1, 0, 2).contiguous())

    r = ShardedRanker(model, rank, world, local_topk=local_topk, merge=merge, min_shard_pois=min_shard, slice_users=slice_users)
    users = SimpleNamespace(offsets=torch.zeros(1), n_users=U, ids=np.arange(U))
    s, i = r.topk(users, k)
    ref_order = [np.lexsort((np.arange(N), -scores[u]))[:k] for u in range(U)]
    ok = s.shape == (U, k) and i.shape == (U, k)
    ok = ok and all(np.array_equal(i[u].numpy(), ref_order[u].astype(np.int32)) for u in range(U))
    ok = ok and all(np.array_equal(s[u].numpy(), scores[u][ref_order[u]]) for u in range(U))
    ok = ok and (r.gc, r.gu) == want_grid and (r.lo, r.hi) == shard_range(N, rank % r.gc, r.gc)
    if want_grid == (2, 1):
        ok = ok and (r.lo, r.hi) == ((0, 512) if rank == 0 else (512, 1000))
    # touched-row exchange of the data-parallel training step: ragged (ids, rows) lists from every rank, rank-major order
    from poi_recommendation_models_b200.distributed import SparseRowExchange
    ex = SparseRowExchange.__new__(SparseRowExchange)
    ex.world, ex.last_bytes = world, 0
    n_mine = 3 + 2 * rank
    ids = torch.arange(n_mine, dtype=torch.int64) * (rank + 1)
    rows = torch.full((n_mine, 4), float(rank + 1))
    gi, gr = ex.exchange(ids, rows)
    want_i = torch.cat([torch.arange(3 + 2 * r, dtype=torch.int64) * (r + 1) for r in range(world)])
    want_r = torch.cat([torch.full((3 + 2 * r, 4), float(r + 1)) for r in range(world)])
    ok = ok and torch.equal(gi, want_i) and torch.equal(gr, want_r)
    # data-parallel gradient averaging
    lin = torch.nn.Linear(4, 3)
    for p_ in lin.parameters():
        p_.grad = torch.full_like(p_, float(rank + 1))
    allreduce_gradients(lin, world)
    mean = (world + 1) / 2.0
    ok = ok and all(torch.allclose(p_.grad, torch.full_like(p_, mean)) for p_ in lin.parameters())
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def _run(world, min_shard, want_grid):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, min_shard, want_grid)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)], res


def test_sharded_ranker_world2_gloo():
    """two catalogue range shards (the layout SURVEY.md §8e prescribes)"""
    _run(2, 0, (2, 1))


def test_user_sliced_ranker_world2_gloo():
    """catalogue too small for two useful shards: the user batch is sliced instead (ragged last slice)"""
    _run(2, 32768, (1, 2))


def test_grid_ranker_world4_gloo():
    """2 catalogue shards x 2 user slices"""
    _run(4, 400, (2, 2))


def test_grid_shape():
    sys.path.insert(0, ROOT)
    from poi_recommendation_models_b200.distributed import grid_shape
    assert grid_shape(40000, 8, 32768) == (1, 8)       # C2: user slices only
    assert grid_shape(1000000, 8, 32768) == (8, 1)     # C4: catalogue shards only
    assert grid_shape(100000, 8, 32768) == (2, 4)
    assert grid_shape(40000, 1, 32768) == (1, 1)
    assert grid_shape(1000000, 6, 32768) == (6, 1)

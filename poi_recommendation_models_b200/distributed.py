"""Multi-GPU plumbing for full-rank evaluation (one process per GPU, `torch.distributed`).

The (user, candidate) pairs are independent — the beta-softmax normalises over the user's *history*, not over
candidates — so the catalogue is split into contiguous POI ranges (and, when the catalogue is too small
for `world` useful shards, the user batch into slices: `grid_shape`); each rank scores its users against its own range
with the fused kernel, keeps a local top-k, and ONE collective per user batch exchanges the lists: an all-gather of [U,k] (fp32 score, int32 global id) = U*k*8*world bytes per rank, followed by an
on-device merge (`nais_topk_merge`, same order rule as the single-GPU top-k, so results are identical to one GPU).

The reference has no distributed code at all; this module is new.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

from . import ops


def shard_range(n: int, rank: int, world: int, align: int = 128) -> Tuple[int, int]:
    """Contiguous POI range of `rank`: boundaries at multiples of `align` (the scorer's candidate tile) so no tile
    straddles two ranks; the last rank takes the remainder."""
    per = (n + world - 1) // world
    per = (per + align - 1) // align * align
    lo = min(n, rank * per)
    hi = min(n, lo + per) if rank < world - 1 else n
    return lo, max(lo, hi)


def gather_lists(score: torch.Tensor, ids: torch.Tensor, world: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """all-gather [U,k] lists from every rank into [U, world, k] (rank-major along dim 1)."""
    U, k = score.shape
    gs = torch.empty(world, U, k, dtype=score.dtype, device=score.device)
    gi = torch.empty(world, U, k, dtype=ids.dtype, device=ids.device)
    # list form: supported by both nccl and gloo (all_gather_into_tensor is nccl-only)
    dist.all_gather(list(gs.unbind(0)), score.contiguous())
    dist.all_gather(list(gi.unbind(0)), ids.contiguous())
    return gs.permute(1, 0, 2).contiguous(), gi.permute(1, 0, 2).contiguous()


def grid_shape(n_items: int, world: int, min_shard_pois: int) -> Tuple[int, int]:
    """(catalogue shards, user slices) with shards * slices == world: the catalogue is split as far as every shard keeps
    at least `min_shard_pois` POIs, the remaining factor splits the users of a batch.  A rank packs the operand image of
    every user it scores, whatever its catalogue range is, so thin shards pay that cost `world` times (C2's 40k POIs on 8
    GPUs: 12 % of a step); a 1M-POI catalogue is sharded 8 ways as SURVEY.md §8e prescribes."""
    gc = 1
    for d in range(1, world + 1):
        if world % d == 0 and (d == 1 or n_items // d >= min_shard_pois):
            gc = d
    return gc, world // gc


class ShardedRanker:
    """predict_topk over a range-sharded catalogue (x user-sliced batches when the catalogue is small, `grid_shape`).
    `local_topk` / `merge` / `slice_users` are injectable so the host-side logic can be exercised with the gloo backend on
    CPU (tests/test_distributed_cpu.py); the defaults are the CUDA ops.  rank = user_slice * catalogue_shards + shard."""

    def __init__(self, model, rank: int = 0, world: int = 1, local_topk: Optional[Callable] = None,
                 merge: Optional[Callable] = None, min_shard_pois: int = 32768, slice_users: Optional[Callable] = None):
        self.model, self.rank, self.world = model, rank, world
        self.n = model.item_num
        self.gc, self.gu = grid_shape(self.n, world, min_shard_pois)
        self.rc, self.ru = rank % self.gc, rank // self.gc
        self.lo, self.hi = shard_range(self.n, self.rc, self.gc)
        self._local = local_topk or self._cuda_local
        self._merge = merge or ops.topk_merge
        self._slice = slice_users or (lambda users, u0, u1: users.slice(u0, u1))
        self.events = []

    def _cuda_local(self, users, k, lo, hi, precision):
        m = self.model
        return ops.fullrank_topk(m.variant, float(m.beta), m._params(), m._catalog, users, k, lo, hi, True, precision)

    def kernel_name(self, precision: str) -> str:
        return {"fp32": "fullrank_fp32_kernel"}.get(precision, "fullrank_tc_kernel")

    @property
    def parallelism(self) -> str:
        return f"{self.gc} catalogue range shard(s) x {self.gu} user slice(s), all-gather of the top-k lists" + (
            " + on-device merge" if self.gc > 1 else "")

    @property
    def last_kernel_ms(self) -> Optional[float]:
        if not self.events:
            return None
        a, b = self.events[-1]
        b.synchronize()
        return a.elapsed_time(b)

    def user_range(self, n_users: int) -> Tuple[int, int, int]:
        """(u0, u1, per): this rank scores users [u0, u1) of a batch; every slice is padded to `per` rows for the gather."""
        per = (n_users + self.gu - 1) // self.gu
        u0 = min(n_users, self.ru * per)
        return u0, min(n_users, u0 + per), per

    def _score_and_exchange(self, users, n_users: int, k: int, precision: str, timed: bool):
        """`users` = this rank's slice.  Local fused scoring + top-k, then ONE all-gather of the lists and (catalogue
        shards > 1) the on-device merge; every rank returns the lists of the whole batch [n_users, k]."""
        u0, u1, per = self.user_range(n_users)
        if timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        s, i = self._local(users, k, self.lo, self.hi, precision)
        if timed:
            b.record()
            self.events = [(a, b)]
        if self.world == 1:
            return s, i
        if s.shape[0] < per:  # last slice: pad with "no entry" rows
            s = torch.cat([s, s.new_full((per - s.shape[0], k), float("-inf"))])
            i = torch.cat([i, i.new_full((per - i.shape[0], k), -1)])
        gs, gi = gather_lists(s, i, self.world)                      # [per, world, k], rank-major along dim 1
        gs = gs.view(per, self.gu, self.gc, k).permute(1, 0, 2, 3).reshape(self.gu * per, self.gc, k)[:n_users]
        gi = gi.view(per, self.gu, self.gc, k).permute(1, 0, 2, 3).reshape(self.gu * per, self.gc, k)[:n_users]
        if self.gc == 1:
            return gs[:, 0].contiguous(), gi[:, 0].contiguous()
        return self._merge(gs.contiguous(), gi.contiguous())

    def topk(self, users, k: int, precision: str = "auto"):
        """`users`: the whole batch (DeviceUsers), resident on every rank."""
        timed = torch.cuda.is_available() and users.offsets.is_cuda
        n_users = getattr(users, "n_users", None)
        if self.gu > 1:
            u0, u1, _ = self.user_range(n_users)
            users = self._slice(users, u0, u1)
        return self._score_and_exchange(users, n_users, k, precision, timed)

    def topk_host(self, indptr: torch.Tensor, items: torch.Tensor, k: int, precision: str = "auto"):
        """End-to-end call: pinned host CSR -> device (only this rank's user slice), rank, top-k lists of the whole batch
        back on the host (sigmoid scores like `forward`, int64 ids)."""
        m = self.model
        dev = next(m.parameters()).device
        n_users = indptr.numel() - 1
        u0, u1, _ = self.user_range(n_users)
        a, b = int(indptr[u0]), int(indptr[u1])
        ip = indptr[u0:u1 + 1].to(dev, non_blocking=True) - a
        it = items[a:b].to(dev, non_blocking=True)
        cat = m._catalog
        users = ops.DeviceUsers(ip, it.to(torch.int32), cat.region[it] if cat.region is not None else None,
                                cat.coords[it].contiguous() if cat.coords is not None else None, u1 - u0, b - a)
        s, i = self._score_and_exchange(users, n_users, k, precision, True)
        s_h = torch.sigmoid(s).to("cpu", non_blocking=True)
        i_h = i.to(torch.int64).to("cpu", non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return s_h, i_h


def allreduce_gradients(model, world: int) -> None:
    """Data-parallel training: average dense gradients across ranks (MLP grads ~17 KB + embedding tables; SURVEY.md
    §8e 'Train partitioning').  One flat bucket -> one NCCL all-reduce."""
    if world <= 1:
        return
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world)
    o = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[o:o + n].view_as(g))
        o += n

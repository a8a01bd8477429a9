"""Multi-GPU plumbing for full-rank evaluation (one process per GPU, `torch.distributed`).

The (user, candidate) pairs are independent — the beta-softmax normalises over the user's *history*, not over
candidates — so the catalogue is split into `world` contiguous POI ranges; each rank scores every user
of the batch against its own range with the fused kernel, keeps a local top-k, and ONE collective per user batch
exchanges the lists: an all-gather of [U,k] (fp32 score, int32 global id) = U*k*8*world bytes per rank, followed by an
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


class ShardedRanker:
    """predict_topk over a range-sharded catalogue.  `local_topk` / `merge` are injectable so the host-side logic can
    be exercised with the gloo backend on CPU (tests/test_distributed_cpu.py); the defaults are the CUDA ops."""

    def __init__(self, model, rank: int = 0, world: int = 1, local_topk: Optional[Callable] = None,
                 merge: Optional[Callable] = None):
        self.model, self.rank, self.world = model, rank, world
        self.n = model.item_num
        self.lo, self.hi = shard_range(self.n, rank, world)
        self._local = local_topk or self._cuda_local
        self._merge = merge or ops.topk_merge
        self.events = []

    def _cuda_local(self, users, k, lo, hi, precision):
        m = self.model
        return ops.fullrank_topk(m.variant, float(m.beta), m._params(), m._catalog, users, k, lo, hi, True, precision)

    def kernel_name(self, precision: str) -> str:
        return {"fp32": "fullrank_fp32_kernel"}.get(precision, "fullrank_tc_kernel")

    @property
    def last_kernel_ms(self) -> Optional[float]:
        if not self.events:
            return None
        a, b = self.events[-1]
        b.synchronize()
        return a.elapsed_time(b)

    def topk(self, users, k: int, precision: str = "auto"):
        timed = torch.cuda.is_available() and users.offsets.is_cuda
        if timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        s, i = self._local(users, k, self.lo, self.hi, precision)
        if timed:
            b.record()
            self.events = [(a, b)]
        if self.world > 1:
            gs, gi = gather_lists(s, i, self.world)
            s, i = self._merge(gs, gi)
        return s, i

    def topk_host(self, indptr: torch.Tensor, items: torch.Tensor, k: int, precision: str = "auto"):
        """End-to-end call: pinned host CSR -> device, rank, top-k lists back on the host (sigmoid scores like
        `forward`, int64 ids)."""
        m = self.model
        dev = next(m.parameters()).device
        ip = indptr.to(dev, non_blocking=True)
        it = items.to(dev, non_blocking=True)
        cat = m._catalog
        users = ops.DeviceUsers(ip, it.to(torch.int32), cat.region[it] if cat.region is not None else None,
                                cat.coords[it].contiguous() if cat.coords is not None else None,
                                indptr.numel() - 1, items.numel())
        s, i = self.topk(users, k, precision)
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

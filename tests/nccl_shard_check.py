"""Launched by tests/test_gpu_plan_and_index.py under torchrun (2 ranks, NCCL): range-sharded / user-sliced ranking through
`ShardedRanker` must equal one GPU bit for bit, and data-parallel training must leave identical parameters on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import nais_testutil as util
    from oracle import nais_oracle as orc
    from poi_recommendation_models_b200 import synthetic
    from poi_recommendation_models_b200.distributed import ShardedRanker

    N, k = 70000, 20
    coords, region, R = synthetic.make_catalog(N, seed=0)
    sd = orc.init_state("region_distance", N, 64, 64, R, 1, seed=1, style="trained")
    m = util.make_model("region_distance", sd, 0.5, device=dev)
    m.set_catalog(region=region, coords=coords)
    rng = np.random.default_rng(2)
    lens = rng.integers(3, 130, 37)  # ragged, odd user count: the last user slice is padded for the gather
    indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    users = m.make_users(indptr, np.concatenate([rng.choice(N, n, replace=False) for n in lens]).astype(np.int64))
    solo = ShardedRanker(m, 0, 1).topk(users, k)
    ok = True
    for grid in ((world, 1), (1, world)):
        r = ShardedRanker(m, rank, world, grid=grid)
        s, i = r.topk(users, k)
        same = bool(torch.equal(i, solo[1]) and torch.equal(s, solo[0]))
        if not same:
            print(f"[rank {rank}] grid {grid}: sharded lists differ from one GPU", file=sys.stderr, flush=True)
        ok = ok and same
        # end to end from pinned host memory as well
        s_h, i_h = r.topk_host(torch.from_numpy(indptr).pin_memory(), torch.from_numpy(users.items.cpu().numpy().astype(np.int64)).pin_memory(), k)
        if not torch.equal(i_h, solo[1].cpu().to(torch.int64)):
            print(f"[rank {rank}] grid {grid}: topk_host lists differ", file=sys.stderr, flush=True)
            ok = False
    # ---- data-parallel training with the touched-row exchange: replicas stay bit-identical and equal the dense global step ----
    from poi_recommendation_models_b200 import batches as PB
    from poi_recommendation_models_b200.distributed import SparseRowExchange
    data = synthetic.make_checkins(16, 5000, seed=3, hist_len=None, max_hist=40, min_hist=3, median_hist=12)
    csr = data.train_csr()
    sd2 = orc.init_state("region_distance", 5000, 64, 64, data.region_num, 1, seed=5, style="trained")
    mt = util.make_model("region_distance", sd2, 0.5, device=dev).train()
    ot = torch.optim.Adagrad(mt.parameters(), lr=0.05)
    ex = SparseRowExchange(mt, ot, world)
    bt = PB.DeviceBatcher(csr, data.region, data.coords, device=dev, seed=0)
    mref = util.make_model("region_distance", sd2, 0.5, device=dev).train()  # every rank replays the GLOBAL step densely
    oref = torch.optim.Adagrad(mref.parameters(), lr=0.05)
    per = 16 // world
    for it in range(10):
        bs = [bt.multi_user_batch(np.arange(r * per, (r + 1) * per), 4, seed=100 * it + r) for r in range(world)]
        mine = bs[rank]
        w_loc = torch.full((mine.B,), 1.0 / mine.B, device=dev)
        ex.step(mine.label, mine, row_weight=w_loc)
        oref.zero_grad()
        tot = 0
        for b_ in bs:  # the global loss: mean over ranks of the per-rank mean BCE
            tot = tot + torch.nn.functional.binary_cross_entropy(torch.sigmoid(mref.segmented_scores(b_)), b_.label) / world
        tot.backward()
        oref.step()
        if it == 0:  # one step: the two paths differ by fp32 summation order only
            for (n_, pa), (_, pb) in zip(mt.named_parameters(), mref.named_parameters()):
                err = float((pa - pb).abs().max()) / max(float(pb.abs().max()), 1e-3)
                if err > 2e-5:
                    print(f"[rank {rank}] {n_}: first sparse-exchange step vs dense global step rel err {err:.2e}", file=sys.stderr, flush=True)
                    ok = False
    for (n_, pa), (_, pb) in zip(mt.named_parameters(), mref.named_parameters()):
        # ten Adagrad steps from a zero accumulator amplify rounding differences (the first updates are +-lr whatever |g| is)
        err = float((pa - pb).abs().max()) / max(float(pb.abs().max()), 1e-3)
        if err > 2e-3:
            print(f"[rank {rank}] {n_}: sparse-exchange step vs dense global step rel err {err:.2e}", file=sys.stderr, flush=True)
            ok = False
        mine_p = pa.detach().clone()
        other = [torch.empty_like(mine_p) for _ in range(world)]
        dist.all_gather(other, mine_p)
        if not all(torch.equal(o_, other[0]) for o_ in other):  # bit-identical replicas after 10 steps
            print(f"[rank {rank}] {n_}: replicas differ", file=sys.stderr, flush=True)
            ok = False
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0 and int(flag.item()) == 1:
        print("NCCL_SHARD_CHECK_OK")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()

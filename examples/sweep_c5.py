#!/usr/bin/env python
"""BASELINE config C5: history length 16..512 x embedding dim 32..256 sweep of the fused full-rank scorer, reporting
users/s, pair-scores/s and the fraction of the measured tensor / FP32 roofline per point.  Under torchrun the (D, H) points
are dealt round-robin to the ranks (a sweep is independent replicas: one point per GPU at a time, no collective on the data
path); rank 0 prints every line.

Tensor-core path (--precision, default tc_auto) for every D = hid of the sweep (16 ... 256).
    python examples/sweep_c5.py [--pois 40000] > profiles/r2_sweep_c5.jsonl
    torchrun --nproc-per-node 8 examples/sweep_c5.py > profiles/r2_sweep_c5.jsonl
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from poi_recommendation_models_b200 import model as M, ops, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pois", type=int, default=40000)
    ap.add_argument("--hist", type=int, nargs="*", default=[16, 32, 64, 128, 256, 512])
    ap.add_argument("--dims", type=int, nargs="*", default=[32, 64, 128, 256])
    ap.add_argument("--precision", default="tc_auto", choices=["tc_auto", "tc_split", "tc_mix", "tc_fast"])
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lines = []
    emit = (lambda d: print(json.dumps(d), flush=True)) if world == 1 else (lambda d: lines.append(d))
    peaks = bench.load_peaks()
    N = args.pois
    coords, region, R = synthetic.make_catalog(N, seed=0)
    point = 0
    for D in args.dims:
        hid = D
        mine = [H for i, H in enumerate(args.hist) if (point + i) % world == rank]
        point += len(args.hist)
        if not mine:
            continue
        torch.manual_seed(1)
        m = M.NAIS_region_distance_Embedding(N, D, hid, 0.5, R, 1)
        with torch.no_grad():
            for name, p in m.named_parameters():
                if name.startswith("embed_"):
                    p.normal_(0, 0.3)
        m = m.to(dev).eval()
        m.set_catalog(region=region, coords=coords)
        prec = args.precision  # every D, hid <= 256 has a tensor tiling (D or hid > 256: ops.resolve_precision falls back to fp32)
        for H in mine:
            F = bench.flops_per_cell(D, hid)
            # size the batch for ~0.1 s per call
            rate = ((3.5e14 if D >= 64 else 1.5e14) if D <= 128 else 2.0e14) if prec != "fp32" else 3.5e13
            users = int(max(148, min(8192, 0.1 * rate / (F * H * N))))
            hist = bench.synth_histories(users, N, H, seed=H)
            indptr = np.arange(0, (users + 1) * H, H, dtype=np.int64)
            u = m.make_users(indptr, hist.reshape(-1))
            try:
                plan = m.ranking_plan(prec)  # per-model constants once (nais_fullrank_prepare), like an evaluation does
                for _ in range(2):
                    ops.fullrank_topk(m.variant, 0.5, m._params(), m._catalog, u, 20, precision=plan.precision, plan=plan)
            except RuntimeError as e:
                emit({"H": H, "D": D, "hid": hid, "precision": prec, "unsupported": str(e)})
                continue
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            a.record()
            for _ in range(reps):
                ops.fullrank_topk(m.variant, 0.5, m._params(), m._catalog, u, 20, precision=plan.precision, plan=plan)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
            tf = users * H * N * F / (ms / 1e3) / 1e12
            peak = peaks["tf_sust"] if prec != "fp32" else 74.0
            emit({"H": H, "D": D, "hid": hid, "pois": N, "users": users, "precision": prec, "ms": ms, "gpu": rank,
                  "users_per_s": users / (ms / 1e3), "pair_scores_per_s": users * N / (ms / 1e3),
                  "alg_tflops": tf, "roofline": "tensor bf16 sustained (measured)" if prec != "fp32" else "fp32 FFMA 74 TFLOP/s (derived)",
                  "frac": tf / peak})
    if world > 1:
        import torch.distributed as dist
        allp = [None] * world
        dist.all_gather_object(allp, lines)
        if rank == 0:
            for d in sorted((x for p_ in allp for x in p_), key=lambda d: (d["D"], d["H"])):
                print(json.dumps(d), flush=True)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""BASELINE config C5: history length 16..512 x embedding dim 32..256 sweep of the fused full-rank scorer, reporting
users/s, pair-scores/s and the fraction of the measured tensor / FP32 roofline per point (one GPU; the 8-GPU run of
bench.py shards the catalogue, per-GPU work is the same kernel).

Tensor-core path (--precision, default tc_auto) where its tiling exists (D, hid <= 128), FP32 CUDA-core path elsewhere.
    python examples/sweep_c5.py [--pois 40000] > profiles/r1_sweep_c5.jsonl
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
    dev = torch.device("cuda:0")
    peaks = bench.load_peaks()
    N = args.pois
    coords, region, R = synthetic.make_catalog(N, seed=0)
    for D in args.dims:
        hid = D
        torch.manual_seed(1)
        m = M.NAIS_region_distance_Embedding(N, D, hid, 0.5, R, 1)
        with torch.no_grad():
            for name, p in m.named_parameters():
                if name.startswith("embed_"):
                    p.normal_(0, 0.3)
        m = m.to(dev).eval()
        m.set_catalog(region=region, coords=coords)
        prec = args.precision if D <= 128 else "fp32"
        for H in args.hist:
            F = bench.flops_per_cell(D, hid)
            # size the batch for ~0.1 s per call
            rate = (3.5e14 if D >= 64 else 1.5e14) if prec != "fp32" else 3.5e13
            users = int(max(148, min(8192, 0.1 * rate / (F * H * N))))
            hist = bench.synth_histories(users, N, H, seed=H)
            indptr = np.arange(0, (users + 1) * H, H, dtype=np.int64)
            u = m.make_users(indptr, hist.reshape(-1))
            try:
                for _ in range(2):
                    ops.fullrank_topk(m.variant, 0.5, m._params(), m._catalog, u, 20, precision=prec)
            except RuntimeError as e:
                print(json.dumps({"H": H, "D": D, "hid": hid, "precision": prec, "unsupported": str(e)}), flush=True)
                continue
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            a.record()
            for _ in range(reps):
                ops.fullrank_topk(m.variant, 0.5, m._params(), m._catalog, u, 20, precision=prec)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
            tf = users * H * N * F / (ms / 1e3) / 1e12
            peak = peaks["tf_sust"] if prec != "fp32" else 74.0
            print(json.dumps({"H": H, "D": D, "hid": hid, "pois": N, "users": users, "precision": prec, "ms": ms,
                              "users_per_s": users / (ms / 1e3), "pair_scores_per_s": users * N / (ms / 1e3),
                              "alg_tflops": tf, "roofline": "tensor bf16 sustained (measured)" if prec != "fp32" else "fp32 FFMA 74 TFLOP/s (derived)",
                              "frac": tf / peak}), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Full-rank throughput of the sibling scorers at the C2 shape (40 000 POIs, history 128, D = hid = 64, top-20):
every model class of model.py through `predict_topk`, FP32 fused kernel vs the tensor-core path.  The two-branch disentangled
model (model.py:410-541) runs one tensor pass per attention branch + the in-kernel haversine bias.

    python examples/bench_siblings.py > profiles/r2_siblings_c2.jsonl
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
from poi_recommendation_models_b200 import model as M, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pois", type=int, default=40000)
    ap.add_argument("--hist", type=int, default=128)
    ap.add_argument("--dim", type=int, default=64)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    N, H, D = args.pois, args.hist, args.dim
    coords, region, R = synthetic.make_catalog(N, seed=0)
    classes = [("basic", lambda: M.NAIS_basic(N, D, D, 0.5)), ("region", lambda: M.NAIS_regionEmbedding(N, D, D, 0.5, R)),
               ("region_distance", lambda: M.NAIS_region_distance_Embedding(N, D, D, 0.5, R, 1)),
               ("distance", lambda: M.NAIS_distance_Embedding(N, D, D, 0.5, R, 1)),
               ("disentangled", lambda: M.NAIS_region_distance_disentangled_Embedding(N, D, D, 0.5, R, 1))]
    for name, make in classes:
        torch.manual_seed(1)
        m = make()
        with torch.no_grad():
            for pn, p in m.named_parameters():
                if pn.startswith("embed_") and "distance" not in pn:
                    p.normal_(0, 0.3)
        m = m.to(dev).eval()
        m.set_catalog(region=region, coords=coords)
        for prec, users in (("fp32", 148), ("tc_auto", 1184)):
            hist = bench.synth_histories(users, N, H, seed=7)
            u = m.make_users(np.arange(0, (users + 1) * H, H, dtype=np.int64), hist.reshape(-1))
            for _ in range(2):
                m.predict_topk(u, 20, precision=prec)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            a.record()
            for _ in range(reps):
                m.predict_topk(u, 20, precision=prec)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
            print(json.dumps({"model": name, "precision": prec, "users_per_call": users, "ms_per_call": round(ms, 3),
                              "users_per_s": round(users / ms * 1000, 1), "pois": N, "hist": H, "D": D, "hid": D}), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""BASELINE config C1 end to end on one B200: synthetic Foursquare-NYC-shaped check-ins (1,083 users x 38,333 POIs,
D = hid = 64, history <= 100), ONE epoch of the reference training loop (run.py:227-255: one user per step, BCE,
Adagrad lr 0.01) on the drop-in module, then full-rank evaluation (validation.py:62-131) with recall@k / NDCG@k —
next to the oracle (CPU restatement of the reference) on a bounded sample:

  * training parity: the first `--ref-train-users` user-steps are replayed by the oracle in float64 from the same
    initial weights and the same `random` stream; parameters are compared after those steps;
  * ranking parity: `--ref-eval-users` users are re-ranked by the oracle with the GPU-trained weights; top-50 lists
    and metrics must agree (outside 1e-4 ties).

    python examples/c1_epoch_and_eval.py [--users 1083] [--ref-train-users 24] [--ref-eval-users 8]
"""
import argparse
import json
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nais_oracle as orc  # checker / CPU arm only
from poi_recommendation_models_b200 import batches as PB, eval_metrics as PM, model as M, synthetic, validation as V


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=1083)
    ap.add_argument("--pois", type=int, default=38333)
    ap.add_argument("--ref-train-users", type=int, default=24)
    ap.add_argument("--ref-eval-users", type=int, default=8)
    ap.add_argument("--precision", default="auto")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    D = hid = 64
    beta, num_ng, lr = 0.5, 4, 0.01
    data = synthetic.make_checkins(args.users, args.pois, seed=0, hist_len=None, max_hist=100, min_hist=5, median_hist=30)
    csr = data.train_csr()
    torch.manual_seed(0)
    random.seed(0)
    np.random.seed(0)
    model = M.NAIS_region_distance_Embedding(args.pois, D, hid, beta, data.region_num, 1).to(dev)
    sd0 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    opt = torch.optim.Adagrad(model.parameters(), lr=lr, weight_decay=0.0)
    order = list(range(args.users))
    random.shuffle(order)
    rng_state = random.getstate()

    # ---- one epoch on the GPU (run.py:227-255) -----------------------------------------------------------------------
    model.train()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loss_sum, snap = 0.0, None
    for step, u in enumerate(order):
        hist, tgt, label, hreg, treg = PB.get_NAIS_batch_region(csr, args.pois, u, num_ng, data.region)
        ll = PB.lat_lon_pairs(data.coords, tgt.cpu().numpy(), hist[0].cpu().numpy())
        opt.zero_grad()
        pred = model(hist, tgt, hreg, treg, ll)
        loss = model.loss_func(pred, label)
        loss.backward()
        opt.step()
        loss_sum += loss.item()
        if step + 1 == args.ref_train_users:
            snap = {k: v.detach().cpu().double().clone() for k, v in model.state_dict().items()}
    torch.cuda.synchronize()
    t_train = time.perf_counter() - t0

    # ---- the same epoch with batches built on the device (f1): same distribution, device RNG -----------------------------
    model2 = M.NAIS_region_distance_Embedding(args.pois, D, hid, beta, data.region_num, 1).to(dev)
    model2.load_state_dict(sd0)
    opt2 = torch.optim.Adagrad(model2.parameters(), lr=lr, weight_decay=0.0)
    bt = PB.DeviceBatcher(csr, data.region, data.coords, device=dev, seed=0)
    model2.train()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loss2 = torch.zeros((), device=dev)
    for u in order:
        hist, tgt, label, hreg, treg, ll = bt.batch(u, num_ng)
        opt2.zero_grad()
        l2 = model2.loss_func(model2(hist, tgt, hreg, treg, ll), label)
        l2.backward()
        opt2.step()
        loss2 += l2.detach()
    torch.cuda.synchronize()
    t_train_dev = time.perf_counter() - t0

    # ---- ... and with the row-sparse Adagrad fused into the backward's segment reduce (f2): no dense table gradients ------
    model3 = M.NAIS_region_distance_Embedding(args.pois, D, hid, beta, data.region_num, 1).to(dev)
    model3.load_state_dict(sd0)
    opt3 = torch.optim.Adagrad(model3.parameters(), lr=lr, weight_decay=0.0)
    bt3 = PB.DeviceBatcher(csr, data.region, data.coords, device=dev, seed=0)
    model3.train()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loss3 = torch.zeros((), device=dev)
    for u in order:
        hist, tgt, label, hreg, treg, ll = bt3.batch(u, num_ng)
        loss3 += model3.fused_adagrad_step(opt3, label, hist, tgt, hreg, treg, ll)
    torch.cuda.synchronize()
    t_train_fused = time.perf_counter() - t0
    fused_vs_dense = max(float((a.detach() - b.detach()).abs().max()) for a, b in zip(model3.parameters(), model2.parameters()))

    # ---- oracle replay of the first steps (float64, same RNG stream) ---------------------------------------------------
    random.setstate(rng_state)
    ref_sd, ref_sum = {k: v.double() for k, v in sd0.items()}, None
    t0 = time.perf_counter()
    for u in order[:args.ref_train_users]:
        h, t, lab, hr, tr = orc.train_batch_region(data.history(u).tolist(), args.pois, num_ng, data.region, random)
        ll = orc.latlon_abs_diff(data.coords, t, h)
        _, ref_sd, ref_sum = orc.train_step_bce(ref_sd, "region_distance", beta, torch.from_numpy(h), torch.from_numpy(t),
                                                torch.from_numpy(hr), torch.from_numpy(tr), torch.from_numpy(ll),
                                                torch.from_numpy(lab), lr, ref_sum, dtype=torch.float64)
    t_ref_train = time.perf_counter() - t0
    train_diff = max(float((snap[k] - ref_sd[k]).abs().max()) for k in ref_sd if k != "embed_distance.weight")

    # ---- full-rank evaluation on the GPU ---------------------------------------------------------------------------------
    k_list = [5, 10, 15, 20, 25, 30]
    ns = argparse.Namespace(topk=50, powerlaw_weight=0.2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res, rec = V.NAIS_region_distance_validation(model, ns, args.users, data.test_positive, data.val_positive, csr,
                                                 data.region, data.coords, k_list, precision=args.precision,
                                                 return_recommended=True)
    torch.cuda.synchronize()
    t_eval = time.perf_counter() - t0
    ndcg10 = PM.ndcg_at_k(data.test_positive, rec, 10)

    # ---- oracle re-ranking of a sample with the GPU-trained weights --------------------------------------------------------
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    cat = orc.Catalog(data.coords, data.region)
    sample = list(range(0, args.users, max(1, args.users // args.ref_eval_users)))[:args.ref_eval_users]
    t0 = time.perf_counter()
    same, valid = 0, 0
    ref_rec = []
    for u in sample:
        r, _, cand, pred = orc.fullrank_user(sd, "region_distance", beta, cat, data.history(u), 50, dtype=torch.float32,
                                             return_all=True)
        ref_rec.append([int(i) for i in r])
        same += int(ref_rec[-1] == rec[u])
        by_id = dict(zip(cand.tolist(), pred.astype(np.float64).tolist()))
        kth = np.sort(pred)[::-1][49]
        valid += int(all(by_id[i] >= kth - 1e-4 * abs(kth) for i in rec[u]))
    t_ref_eval = time.perf_counter() - t0
    sub_pos = [data.test_positive[u] for u in sample]
    m_gpu = [PM.recall_at_k(sub_pos, [rec[u] for u in sample], k) for k in k_list]
    m_ref = [orc.recall_at_k(sub_pos, ref_rec, k) for k in k_list]

    report = {
        "config": f"C1: {args.users} users x {args.pois} POIs, D=hid=64, H<=100, 1 epoch BCE/Adagrad + full-rank eval",
        "train": {"gpu_s": t_train, "gpu_users_per_s": args.users / t_train, "epoch_loss_sum": loss_sum,
                  "device_batcher_gpu_s": t_train_dev, "device_batcher_users_per_s": args.users / t_train_dev,
                  "device_batcher_epoch_loss_sum": float(loss2),
                  "fused_adagrad_gpu_s": t_train_fused, "fused_adagrad_users_per_s": args.users / t_train_fused,
                  "fused_adagrad_epoch_loss_sum": float(loss3), "fused_vs_dense_max_abs_param_diff_after_epoch": fused_vs_dense,
                  "oracle_users_per_s": args.ref_train_users / t_ref_train, "oracle_users": args.ref_train_users,
                  "max_abs_param_diff_after_oracle_steps": train_diff},
        "eval": {"gpu_s": t_eval, "gpu_users_per_s": args.users / t_eval, "precision": args.precision,
                 "recall_test@k": dict(zip(map(str, k_list), res[4])), "recall_val@k": dict(zip(map(str, k_list), res[1])),
                 "ndcg_test@10": ndcg10,
                 "oracle_users": len(sample), "oracle_users_per_s": len(sample) / t_ref_eval,
                 "identical_top50_lists": same, "valid_top50_lists_within_1e-4_ties": valid,
                 "recall@k_sample_gpu": m_gpu, "recall@k_sample_oracle": m_ref, "recall_identical": m_gpu == m_ref},
        "host_cores": os.cpu_count(),
    }
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()

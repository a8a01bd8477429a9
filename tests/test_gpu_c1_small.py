"""BASELINE config C1 in miniature, end to end through the drop-in API: the reference's training schedule (run.py:227-255 — one
user per step, `get_NAIS_batch_region` under Python's `random`, BCE, Adagrad) replayed step by step by the float64 oracle, then the
reference's validator (validation.py:62-131) on the trained weights against the oracle's ranking: lists, recall@k, precision@k."""
import argparse
import random

import numpy as np
import pytest
import torch

from oracle import nais_oracle as orc
from poi_recommendation_models_b200 import batches as PB, eval_metrics as PM, model as M, synthetic, validation as V

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("pairs_precision", ["auto", "fp32"])
def test_c1_epoch_then_eval_matches_the_oracle(pairs_precision):
    U, N, D, hid, beta, num_ng, lr = 48, 2500, 64, 64, 0.5, 4, 0.01
    data = synthetic.make_checkins(U, N, seed=2, hist_len=None, max_hist=60, min_hist=4, median_hist=18)
    csr = data.train_csr()
    torch.manual_seed(0)
    random.seed(0)
    model = M.NAIS_region_distance_Embedding(N, D, hid, beta, data.region_num, 1).cuda().train()
    model.pairs_precision = pairs_precision
    sd0 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    opt = torch.optim.Adagrad(model.parameters(), lr=lr, weight_decay=0.0)
    order = list(range(U))
    random.shuffle(order)
    state = random.getstate()
    # ---- one epoch on the GPU, the reference's loop body (run.py:235-254) ------------------------------------------------
    for u in order:
        hist, tgt, label, hreg, treg = PB.get_NAIS_batch_region(csr, N, u, num_ng, data.region)
        ll = PB.lat_lon_pairs(data.coords, tgt.cpu().numpy(), hist[0].cpu().numpy())
        opt.zero_grad()
        loss = model.loss_func(model(hist, tgt, hreg, treg, ll), label)
        loss.backward()
        opt.step()
    # ---- the same epoch by the oracle in float64, same RNG stream -----------------------------------------------------------
    random.setstate(state)
    ref_sd, ref_sum = {k: v.double() for k, v in sd0.items()}, None
    for u in order:
        h, t, lab, hr, tr = orc.train_batch_region(data.history(u).tolist(), N, num_ng, data.region, random)
        ll = orc.latlon_abs_diff(data.coords, t, h)
        _, ref_sd, ref_sum = orc.train_step_bce(ref_sd, "region_distance", beta, torch.from_numpy(h), torch.from_numpy(t), torch.from_numpy(hr),
                                                torch.from_numpy(tr), torch.from_numpy(ll), torch.from_numpy(lab), lr, ref_sum, dtype=torch.float64)
    got = {k: v.detach().cpu().double() for k, v in model.state_dict().items()}
    for k in ref_sd:
        if k == "embed_distance.weight":
            continue  # allocated, never read (model.py:204)
        # Adagrad from a zero accumulator turns the SIGN of a gradient element into a +-lr step whatever its size, so an element
        # whose gradient is below the kernels' rounding (3e-7 of the tensor's maximum for the FP32 kernels, ~3e-6 for the tcgen05
        # ones; single-step gradient parity is pinned at 2e-4 in test_gpu_backward / test_gpu_pairs_tc) may land a whole step
        # away: the bulk must agree tightly, outliers must be rare and never exceed a couple of steps
        diff = (got[k] - ref_sd[k]).abs()
        assert float(diff.mean()) <= 2e-5, (k, float(diff.mean()))
        assert float((diff > 1e-3).double().mean()) <= 2e-3, (k, float((diff > 1e-3).double().mean()))
        assert float(diff.max()) <= 4 * lr, (k, float(diff.max()))
    # ---- full-rank evaluation of every user with the GPU-trained weights (validation.py:62-131) -----------------------------
    k_list = [5, 10, 15, 20, 25, 30]
    ns = argparse.Namespace(topk=50, powerlaw_weight=0.2)
    res, rec = V.NAIS_region_distance_validation(model, ns, U, data.test_positive, data.val_positive, csr, data.region, data.coords,
                                                 k_list, return_recommended=True)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    cat = orc.Catalog(data.coords, data.region)
    ref_rec = []
    for u in range(U):
        r, _, cand, pred = orc.fullrank_user(sd, "region_distance", beta, cat, data.history(u), 50, dtype=torch.float64, return_all=True)
        ref_rec.append([int(i) for i in r])
        by_id = dict(zip(cand.tolist(), pred.tolist()))
        kth = np.sort(pred)[::-1][49]
        assert all(by_id[i] >= kth - 1e-4 * abs(kth) for i in rec[u]), u  # a valid top-50 of the oracle's scores
    same = sum(int(a == b) for a, b in zip(rec, ref_rec))
    assert same >= U - 2, same  # (after one epoch from the 0.01 init scores sit near 0.5: a near-tie may swap two neighbours)
    for i, k in enumerate(k_list):
        assert abs(res[4][i] - orc.recall_at_k(data.test_positive, ref_rec, k)) <= 2.0 / U
        assert abs(res[3][i] - orc.precision_at_k(data.test_positive, ref_rec, k)) <= 2.0 / (U * k) + 1e-12
        assert res[4][i] == PM.recall_at_k(data.test_positive, rec, k)  # the validator's numbers are eval_metrics' numbers

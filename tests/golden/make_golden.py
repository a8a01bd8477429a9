"""Generate the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE.

    python tests/golden/make_golden.py            # needs /root/reference (not present on the GPU box)

The reference has no tests or golden vectors of its own, so parity is pinned on outputs of the
reference code itself: every `.npz` here holds seeded inputs, the parameter set and what the reference returned.
The fixtures travel with the repo; the reference does not.

Fixtures
  scorer_<variant>.npz   model.<class>.attention_network / forward on explicit pair batches (fp32 and .double()),
                         plus BCE-loss gradients of every parameter for the trained-like state
  validation_rd.npz      validation.NAIS_region_distance_validation on a tiny dataset: recommended_list (top-50),
                         metrics from eval_metrics.{precision,recall,hitrate}_at_k
  batches.npz            batches.get_NAIS_batch_region under random.seed(123)
  geo.npz                powerLaw.dist and run.lat_lon_mat-style |dlat|,|dlon| values
"""
from __future__ import annotations

import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import nais_oracle as orc  # noqa: E402
from oracle import ref_shim  # noqa: E402
from poi_recommendation_models_b200 import synthetic  # noqa: E402

torch.set_num_threads(4)


def _ref_model(ref_model, variant, N, D, hid, beta, R, sd):
    cls = getattr(ref_model, orc.VARIANTS[variant]["cls"])
    if variant == "basic":
        m = cls(N, D, hid, beta)
    elif variant == "region":
        m = cls(N, D, hid, beta, R)
    else:
        m = cls(N, D, hid, beta, R, 1)
    m.DEVICE = torch.device("cpu")
    missing = m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m.eval()


def _call(m, variant, hist, tgt, hreg, treg, aux, pre_sigmoid=True):
    fn = m.attention_network if pre_sigmoid else m.forward
    if variant == "basic":
        return fn(hist, tgt)
    if variant == "region":
        return fn(hist, tgt, hreg, treg)
    if variant == "distance":
        return fn(hist, tgt, aux) if pre_sigmoid else fn(hist, tgt, hreg, treg, aux)
    return fn(hist, tgt, hreg, treg, aux)


def make_scorer(ref_model):
    for variant in orc.VARIANTS:
        rng = np.random.default_rng(7)
        N, hid, beta = 400, 64 if variant == "region_distance" else 32, 0.5
        D = 64 if variant == "region_distance" else 32
        coords, region, R = synthetic.make_catalog(N, seed=3)
        out = {"variant": variant, "N": N, "D": D, "hid": hid, "beta": beta, "R": R}
        for style in ("reference", "trained"):
            sd = orc.init_state(variant, N, D, hid, R, 1, seed=11, style=style)
            B, H = 40, 13
            hist = np.stack([rng.choice(N, H, replace=False) for _ in range(B)]).astype(np.int64)
            tgt = rng.integers(0, N, B).astype(np.int64)
            tgt[::5] = hist[::5, 2]  # live "history item is the target" mask (training positives)
            hreg, treg = region[hist], region[tgt]
            if orc.VARIANTS[variant]["dist"] == "latlon":
                aux = orc.latlon_abs_diff(coords, tgt, hist)
            elif orc.VARIANTS[variant]["dist"] == "km":
                aux = orc.dist_km(coords[tgt][:, None, 0], coords[tgt][:, None, 1], coords[hist][:, :, 0],
                                  coords[hist][:, :, 1]).astype(np.float32)
            else:
                aux = np.zeros((B, H), dtype=np.float32)
            t = lambda a: torch.from_numpy(a)
            m = _ref_model(ref_model, variant, N, D, hid, beta, R, sd)
            with torch.no_grad():
                s32 = _call(m, variant, t(hist), t(tgt), t(hreg), t(treg), t(aux)).numpy()
                f32 = _call(m, variant, t(hist), t(tgt), t(hreg), t(treg), t(aux), pre_sigmoid=False).numpy()
                m64 = _ref_model(ref_model, variant, N, D, hid, beta, R, sd).double()
                s64 = _call(m64, variant, t(hist), t(tgt), t(hreg), t(treg), t(aux).double()).numpy()
            pre = style + "_"
            out.update({pre + "hist": hist, pre + "tgt": tgt, pre + "hreg": hreg, pre + "treg": treg, pre + "aux": aux,
                        pre + "score32": s32, pre + "forward32": f32, pre + "score64": s64})
            for k, v in sd.items():
                out[pre + "sd." + k] = v.numpy()
            if style == "trained":
                # gradients of the reference training loss (run.py:248-253) in float64
                m64.train()
                m64.zero_grad()
                label = torch.from_numpy((rng.random(B) < 0.3).astype(np.float64))
                pred = _call(m64, variant, t(hist), t(tgt), t(hreg), t(treg), t(aux).double(), pre_sigmoid=False)
                if variant in ("basic", "region"):
                    m64.eval()  # dropout(0.5) in train mode is random (model.py:71,162); grads are pinned in eval mode
                    pred = _call(m64, variant, t(hist), t(tgt), t(hreg), t(treg), t(aux).double(), pre_sigmoid=False)
                loss = m64.loss_func(pred, label)
                loss.backward()
                out["grad_label"] = label.numpy()
                out["grad_loss"] = np.float64(loss.item())
                for k, p in m64.named_parameters():
                    out["grad." + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
        np.savez_compressed(os.path.join(HERE, f"scorer_{variant}.npz"), **out)
        print("wrote scorer", variant)


def make_validation(ref_model, ref_validation, ref_metrics):
    U, N, D, hid, beta = 6, 500, 64, 64, 0.5
    data = synthetic.make_checkins(U, N, hist_len=None, seed=5, max_hist=40, min_hist=5, median_hist=15)
    sd = orc.init_state("region_distance", N, D, hid, data.region_num, 1, seed=21, style="trained")
    m = _ref_model(ref_model, "region_distance", N, D, hid, beta, data.region_num, sd)
    latlon_mat = np.abs(data.coords[:, None, :] - data.coords[None, :, :])  # run.py:47-54, vectorised
    captured = {}

    def capture(positive, recommended, k_list):
        captured.setdefault("rec", [list(r) for r in recommended])
        return ([ref_metrics.precision_at_k(positive, recommended, k) for k in k_list],
                [ref_metrics.recall_at_k(positive, recommended, k) for k in k_list],
                [ref_metrics.hitrate_at_k(positive, recommended, k) for k in k_list])

    ref_validation.eval_metrics = types.SimpleNamespace(evaluate_mp=capture)
    args = types.SimpleNamespace(powerlaw_weight=0.2, topk=50)
    k_list = [5, 10, 15, 20, 25, 30]
    with torch.no_grad():
        res = ref_validation.NAIS_region_distance_validation(m, args, U, data.test_positive, data.val_positive,
                                                             data.train_csr(), data.region, latlon_mat, k_list)
    ref_validation.eval_metrics = ref_metrics
    out = {"U": U, "N": N, "D": D, "hid": hid, "beta": beta, "seed": 5, "k_list": np.array(k_list),
           "rec": np.array(captured["rec"], dtype=np.int64), "metrics": np.array(res, dtype=np.float64),
           "coords": data.coords, "region": data.region, "indptr": data.indptr, "indices": data.indices,
           "val_flat": np.concatenate([np.array(v, dtype=np.int64) for v in data.val_positive]),
           "val_ptr": np.cumsum([0] + [len(v) for v in data.val_positive]),
           "test_flat": np.concatenate([np.array(v, dtype=np.int64) for v in data.test_positive]),
           "test_ptr": np.cumsum([0] + [len(v) for v in data.test_positive])}
    for k, v in sd.items():
        out["sd." + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "validation_rd.npz"), **out)
    print("wrote validation_rd; recall@10 val/test:", res[1][1], res[4][1])


def make_batches(ref_batches):
    data = synthetic.make_checkins(4, 300, hist_len=None, seed=9, max_hist=20, min_hist=5, median_hist=10)
    out = {"indptr": data.indptr, "indices": data.indices, "region": data.region, "N": 300, "num_ng": 4}
    csr = data.train_csr()
    for u in range(4):
        random.seed(123 + u)
        hist, tgt, label, hreg, treg = ref_batches.get_NAIS_batch_region(csr, 300, u, 4, data.region)
        out.update({f"u{u}.hist": hist.numpy(), f"u{u}.tgt": tgt.numpy(), f"u{u}.label": label.numpy(),
                    f"u{u}.hreg": hreg.numpy(), f"u{u}.treg": treg.numpy()})
        h2, t2, _, hr2, tr2 = ref_batches.get_NAIS_batch_test_region(csr, u, data.region)
        out.update({f"u{u}.test_tgt": t2.numpy(), f"u{u}.test_hist0": h2[0].numpy()})
    np.savez_compressed(os.path.join(HERE, "batches.npz"), **out)
    print("wrote batches")


def make_geo(ref_powerlaw):
    rng = np.random.default_rng(1)
    a = np.stack([40.5 + rng.random(64) * 0.5, -74.1 + rng.random(64) * 0.5], 1)
    b = a.copy()
    b[10:] = np.stack([40.5 + rng.random(54) * 0.5, -74.1 + rng.random(54) * 0.5], 1)
    b[10:20] = a[10:20] + rng.normal(0, 3e-7, (10, 2))  # inside / around the 1e-6 short-circuit
    d = np.array([ref_powerlaw.dist(tuple(x), tuple(y)) for x, y in zip(a, b)])
    np.savez_compressed(os.path.join(HERE, "geo.npz"), a=a, b=b, dist=d, absdiff=np.abs(a - b))
    print("wrote geo")


if __name__ == "__main__":
    ref_model, ref_batches, ref_validation, ref_metrics, ref_powerlaw = ref_shim.load_reference(
        "model", "batches", "validation", "eval_metrics", "powerLaw")
    make_scorer(ref_model)
    make_validation(ref_model, ref_validation, ref_metrics)
    make_batches(ref_batches)
    make_geo(ref_powerlaw)

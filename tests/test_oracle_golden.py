"""The oracle (oracle/nais_oracle.py) against outputs of the unmodified reference (tests/golden/*.npz, written by
tests/golden/make_golden.py), and against the live reference when /root/reference exists."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import nais_oracle as orc
from oracle import ref_shim
from poi_recommendation_models_b200 import synthetic

VARIANTS = list(orc.VARIANTS)


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _sd(z, prefix):
    return {k[len(prefix):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix)}


def _cond_err(got, ref64, scale):
    return float(np.max(np.abs(got - ref64) / np.maximum(np.abs(ref64), scale)))


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("style", ["reference", "trained"])
def test_scorer_matches_reference(golden_dir, variant, style):
    z = _load(golden_dir, f"scorer_{variant}.npz")
    sd = _sd(z, style + "_sd.")
    t = lambda k: torch.from_numpy(z[style + "_" + k])
    beta = float(z["beta"])
    s32 = orc.attention_network(sd, variant, beta, t("hist"), t("tgt"), t("hreg"), t("treg"), t("aux")).numpy()
    s64 = orc.attention_network(sd, variant, beta, t("hist"), t("tgt"), t("hreg"), t("treg"), t("aux"),
                                dtype=torch.float64).numpy()
    f32 = orc.forward(sd, variant, beta, t("hist"), t("tgt"), t("hreg"), t("treg"), t("aux")).numpy()
    # fp64 restatement == fp64 reference to rounding; fp32 within the reference's own fp32-vs-fp64 gap
    np.testing.assert_allclose(s64, z[style + "_score64"], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(s32, z[style + "_score32"], rtol=2e-5, atol=2e-7)
    np.testing.assert_allclose(f32, z[style + "_forward32"], rtol=1e-6, atol=0)


@pytest.mark.parametrize("variant", VARIANTS)
def test_gradients_match_reference(golden_dir, variant):
    z = _load(golden_dir, f"scorer_{variant}.npz")
    sd = _sd(z, "trained_sd.")
    t = lambda k: torch.from_numpy(z["trained_" + k])
    label = torch.from_numpy(z["grad_label"])
    P = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    pred = torch.sigmoid(orc.attention_network(P, variant, float(z["beta"]), t("hist"), t("tgt"), t("hreg"), t("treg"),
                                               t("aux"), dtype=torch.float64))
    loss = orc.bce_loss(pred, label)
    loss.backward()
    assert abs(loss.item() - float(z["grad_loss"])) < 1e-12
    for k, p in P.items():
        ref = z["grad." + k]
        got = p.grad.numpy() if p.grad is not None else np.zeros_like(ref)
        np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-14, err_msg=k)


def test_validation_flow_matches_reference(golden_dir):
    z = _load(golden_dir, "validation_rd.npz")
    sd = _sd(z, "sd.")
    cat = orc.Catalog(z["coords"], z["region"])
    rec = orc.fullrank(sd, "region_distance", float(z["beta"]), cat, z["indptr"], z["indices"], topk=50)
    assert np.array_equal(np.array(rec), z["rec"])
    val = [z["val_flat"][z["val_ptr"][u]:z["val_ptr"][u + 1]].tolist() for u in range(int(z["U"]))]
    test = [z["test_flat"][z["test_ptr"][u]:z["test_ptr"][u + 1]].tolist() for u in range(int(z["U"]))]
    k_list = z["k_list"].tolist()
    pv, rv, hv = orc.evaluate(val, rec, k_list)
    pt, rt, ht = orc.evaluate(test, rec, k_list)
    assert np.array_equal(np.array([pv, rv, hv, pt, rt, ht]), z["metrics"])  # bit-identical python floats


@pytest.mark.parametrize("variant", ["basic", "region"])
def test_sibling_validators_match_reference(golden_dir, variant):
    """validation.NAIS_validation (validation.py:7-31, chunks of 1024) and NAIS_region_validation (:34-59): the oracle's full-rank
    flow gives the reference's recommended lists and metrics for the sibling scorers too."""
    z = _load(golden_dir, f"validation_{variant}.npz")
    sd = _sd(z, "sd.")
    cat = orc.Catalog(z["coords"], z["region"])
    rec = orc.fullrank(sd, variant, float(z["beta"]), cat, z["indptr"], z["indices"], topk=50, chunk=1024)
    assert np.array_equal(np.array(rec), z["rec"])
    val = [z["val_flat"][z["val_ptr"][u]:z["val_ptr"][u + 1]].tolist() for u in range(int(z["U"]))]
    test = [z["test_flat"][z["test_ptr"][u]:z["test_ptr"][u + 1]].tolist() for u in range(int(z["U"]))]
    k_list = z["k_list"].tolist()
    pv, rv, hv = orc.evaluate(val, rec, k_list)
    pt, rt, ht = orc.evaluate(test, rec, k_list)
    assert np.array_equal(np.array([pv, rv, hv, pt, rt, ht]), z["metrics"])


def test_batches_match_reference(golden_dir):
    z = _load(golden_dir, "batches.npz")
    for u in range(4):
        random.seed(123 + u)
        hist, tgt, label, hreg, treg = orc.train_batch_region(
            z["indices"][z["indptr"][u]:z["indptr"][u + 1]].tolist(), int(z["N"]), int(z["num_ng"]), z["region"], random)
        for name, got in (("hist", hist), ("tgt", tgt), ("label", label), ("hreg", hreg), ("treg", treg)):
            assert np.array_equal(got, z[f"u{u}.{name}"]), (u, name)
        cand = orc.test_candidates(z["indices"][z["indptr"][u]:z["indptr"][u + 1]], int(z["N"]))
        assert np.array_equal(cand, z[f"u{u}.test_tgt"])


def test_geo_matches_reference(golden_dir):
    z = _load(golden_dir, "geo.npz")
    a, b = z["a"], z["b"]
    np.testing.assert_allclose(orc.dist_km(a[:, 0], a[:, 1], b[:, 0], b[:, 1]), z["dist"], rtol=1e-12, atol=1e-9)
    got = orc.latlon_abs_diff(np.concatenate([a, b]), np.arange(64), (np.arange(64) + 64)[:, None])
    assert np.array_equal(got[:, 0, :], z["absdiff"].astype(np.float32))


def test_rank_metrics_match_reference(golden_dir):
    """precision_at_k_per_sample / apk / mapk (eval_metrics.py:29-34, 70-125): oracle == reference outputs, bit for bit."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk_rank", os.path.join(golden_dir, "make_golden_rank_metrics.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    g = np.load(os.path.join(golden_dir, "rank_metrics.npz"))
    actual, predicted = mk.make_lists()
    for j, k in enumerate(g["ks"].tolist()):
        for u, (a, p) in enumerate(zip(actual, predicted)):
            assert orc.apk(a, p, k) == g["apk"][u, j]
            assert orc.precision_at_k_per_sample(a, p, k) == g["pps"][u, j]
        assert orc.mapk(actual, predicted, k) == g["mapk"][j]


def test_metric_edge_cases():
    actual = [[1, 2], [], [5]]
    pred = [[2, 9, 1], [3, 4, 5], [6, 7, 8]]
    assert orc.precision_at_k(actual, pred, 2) == (0.5 + 0 + 0) / 3
    assert orc.recall_at_k(actual, pred, 2) == (0.5 + 0.0) / 2  # users without positives are skipped
    assert orc.hitrate_at_k(actual, pred, 3) == 0.5
    assert 0 < orc.ndcg_at_k(actual, pred, 3) < 1


def test_all_masked_row_is_nan():
    # H=1 and the target is that item: 0/0^beta (SURVEY.md §3.3 quirk ii)
    sd = orc.init_state("basic", 5, 8, 8, seed=0, style="trained")
    s = orc.attention_network(sd, "basic", 0.5, torch.tensor([[3]]), torch.tensor([3]))
    assert torch.isnan(s).all()


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("variant", VARIANTS)
def test_live_reference_random_states(variant):
    """Fresh seeds against the live, unmodified reference classes."""
    ref_model = ref_shim.load_reference("model")
    rng = np.random.default_rng(99)
    N, D, hid, beta = 200, 16, 24, 0.3
    coords, region, R = synthetic.make_catalog(N, seed=8)
    sd = orc.init_state(variant, N, D, hid, R, 1, seed=5, style="trained")
    cls = getattr(ref_model, orc.VARIANTS[variant]["cls"])
    m = cls(N, D, hid, beta) if variant == "basic" else cls(N, D, hid, beta, R) if variant == "region" else cls(N, D, hid, beta, R, 1)
    m.DEVICE = torch.device("cpu")
    m.load_state_dict(sd)
    m.eval().double()
    B, H = 17, 9
    hist = torch.from_numpy(np.stack([rng.choice(N, H, replace=False) for _ in range(B)]))
    tgt = torch.from_numpy(rng.integers(0, N, B))
    hreg, treg = torch.from_numpy(region)[hist], torch.from_numpy(region)[tgt]
    kind = orc.VARIANTS[variant]["dist"]
    if kind == "latlon":
        aux = torch.from_numpy(orc.latlon_abs_diff(coords, tgt.numpy(), hist.numpy()))
    else:
        aux = torch.rand(B, H) * 5
    with torch.no_grad():
        if variant == "basic":
            ref = m.attention_network(hist, tgt)
        elif variant == "region":
            ref = m.attention_network(hist, tgt, hreg, treg)
        elif variant == "distance":
            ref = m.attention_network(hist, tgt, aux.double())
        else:
            ref = m.attention_network(hist, tgt, hreg, treg, aux.double())
    got = orc.attention_network(sd, variant, beta, hist, tgt, hreg, treg, aux, dtype=torch.float64)
    np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=1e-11, atol=1e-13)

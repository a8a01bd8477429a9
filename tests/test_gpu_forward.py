"""GPU parity (through the C ABI): pair scorer forward, fused full-rank scores, top-k, merge — against the oracle and
the golden fixtures written by the unmodified reference."""
import numpy as np
import pytest
import torch

from oracle import nais_oracle as orc
from poi_recommendation_models_b200 import ops, synthetic
import nais_testutil as util

pytestmark = pytest.mark.gpu
VARIANTS = list(orc.VARIANTS)


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("style", ["reference", "trained"])
def test_pairs_forward_matches_reference_golden(variant, style):
    z = util.load_golden(f"scorer_{variant}.npz")
    sd = util.golden_sd(z, style + "_sd.")
    beta = float(z["beta"])
    m = util.make_model(variant, sd, beta)
    a = {k: z[f"{style}_{k}"] for k in ("hist", "tgt", "hreg", "treg", "aux")}
    t = {k: _dev(v) for k, v in a.items()}
    with torch.no_grad():
        s = util.call(m, variant, t["hist"], t["tgt"], t["hreg"], t["treg"], t["aux"]).cpu().numpy()
        f = util.call(m, variant, t["hist"], t["tgt"], t["hreg"], t["treg"], t["aux"], pre_sigmoid=False).cpu().numpy()
    _, scale = orc.attention_network_with_scale(sd, variant, beta, *(torch.from_numpy(a[k]) for k in ("hist", "tgt", "hreg", "treg", "aux")),
                                                dtype=torch.float64)
    ref64 = z[f"{style}_score64"]
    assert util.cond_err(s, ref64, scale.numpy()) < util.TOL
    np.testing.assert_allclose(f, z[f"{style}_forward32"], rtol=util.TOL, atol=0)  # what forward() returns / is ranked
    # and as close to fp64 truth as the reference's own fp32 path is (x4 slack)
    ref_gap = util.cond_err(z[f"{style}_score32"], ref64, scale.numpy())
    assert util.cond_err(s, ref64, scale.numpy()) <= max(4 * ref_gap, 2e-6)


@pytest.mark.parametrize("H", [1, 3, 13, 100, 128, 129, 300])
def test_pairs_forward_history_lengths(H):
    rng = np.random.default_rng(H)
    N, D, hid, beta, B = 900, 64, 64, 0.5, 37
    coords, region, R = synthetic.make_catalog(N, seed=1)
    sd = orc.init_state("region_distance", N, D, hid, R, 1, seed=3, style="trained")
    hist = np.stack([rng.choice(N, H, replace=False) for _ in range(B)]).astype(np.int64)
    tgt = rng.integers(0, N, B).astype(np.int64)
    if H > 1:
        tgt[::4] = hist[::4, H // 2]
    aux = orc.latlon_abs_diff(coords, tgt, hist)
    m = util.make_model("region_distance", sd, beta)
    with torch.no_grad():
        s = m.attention_network(_dev(hist), _dev(tgt), _dev(region[hist]), _dev(region[tgt]), _dev(aux)).cpu().numpy()
    ref, scale = orc.attention_network_with_scale(sd, "region_distance", beta, torch.from_numpy(hist), torch.from_numpy(tgt),
                                                  torch.from_numpy(region[hist]), torch.from_numpy(region[tgt]),
                                                  torch.from_numpy(aux), dtype=torch.float64)
    assert util.cond_err(s, ref.numpy(), scale.numpy()) < util.TOL


@pytest.mark.parametrize("D,hid", [(32, 32), (64, 48), (128, 128), (256, 64), (64, 200)])
def test_pairs_forward_shapes(D, hid):
    rng = np.random.default_rng(D + hid)
    N, beta, B, H = 500, 0.7, 21, 17
    coords, region, R = synthetic.make_catalog(N, seed=2)
    sd = orc.init_state("region_distance", N, D, hid, R, 1, seed=4, style="trained")
    hist = np.stack([rng.choice(N, H, replace=False) for _ in range(B)]).astype(np.int64)
    tgt = rng.integers(0, N, B).astype(np.int64)
    aux = orc.latlon_abs_diff(coords, tgt, hist)
    m = util.make_model("region_distance", sd, beta)
    with torch.no_grad():
        s = m.attention_network(_dev(hist), _dev(tgt), _dev(region[hist]), _dev(region[tgt]), _dev(aux)).cpu().numpy()
    ref, scale = orc.attention_network_with_scale(sd, "region_distance", beta, torch.from_numpy(hist), torch.from_numpy(tgt),
                                                  torch.from_numpy(region[hist]), torch.from_numpy(region[tgt]),
                                                  torch.from_numpy(aux), dtype=torch.float64)
    assert util.cond_err(s, ref.numpy(), scale.numpy()) < util.TOL


def test_all_masked_row_is_nan_like_reference():
    sd = orc.init_state("basic", 50, 32, 32, seed=0, style="trained")
    m = util.make_model("basic", sd, 0.5)
    with torch.no_grad():
        s = m.attention_network(torch.tensor([[3], [4]]).cuda(), torch.tensor([3, 5]).cuda()).cpu()
    assert torch.isnan(s[0]) and torch.isfinite(s[1])


def test_cpu_tensors_fail_loudly():
    sd = orc.init_state("basic", 50, 32, 32, seed=0)
    m = util.make_model("basic", sd, 0.5, device="cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        m.attention_network(torch.tensor([[3]]), torch.tensor([4]))


def _fullrank_case(variant, U, N, seed, D=64, hid=64, beta=0.5, **kw):
    data = synthetic.make_checkins(U, N, seed=seed, **kw)
    sd = orc.init_state(variant, N, D, hid, data.region_num, 1, seed=seed + 1, style="trained")
    m = util.make_model(variant, sd, beta)
    m.set_catalog(region=data.region, coords=data.coords)
    return data, sd, m


PREC_TOL = {"fp32": util.TOL, "tc_split": util.TOL, "tc_mix": util.TOL, "tc_auto": util.TOL, "tc_fast": 5e-4}  # tc_fast: single-pass fp16 logits (documented)


@pytest.mark.parametrize("variant", ["region_distance", "region", "basic", "distance"])
@pytest.mark.parametrize("precision", ["fp32", "tc_split", "tc_mix", "tc_fast"])
def test_fullrank_scores_match_oracle(variant, precision):
    U, N, beta = 5, 700, 0.5
    data, sd, m = _fullrank_case(variant, U, N, seed=11, hist_len=None, max_hist=45, min_hist=2, median_hist=12)
    users = m.make_users(data.indptr, data.indices)
    got = ops.fullrank_scores(variant, beta, m._params(), m._catalog, users, precision=precision).cpu().numpy()
    for u in range(U):
        ref, scale = util.oracle_user_scores(sd, variant, beta, data.coords, data.region, data.history(u), np.arange(N))
        assert util.cond_err(got[u], ref, scale) < PREC_TOL[precision], (u, util.cond_err(got[u], ref, scale))


@pytest.mark.parametrize("precision", ["fp32", "tc_split", "tc_mix", "tc_auto"])
@pytest.mark.parametrize("D,hid", [(32, 32), (64, 64), (64, 128)])
def test_fullrank_disentangled_fused_haversine(precision, D, hid):
    """Two-branch disentangled model (model.py:467-534) through the fused full-rank paths: dist_km (powerLaw.dist, law of cosines
    in float64) is formed in-kernel as a haversine of centred fp32 coordinates.  FP32: one kernel loops over both branches.
    Tensor path: one scoring pass per branch, the second adds the first's scores before it ranks (csrc/nais_tc.cu
    fullrank_tc_run)."""
    U, N, beta = 6, 1000, 0.5
    data, sd, m = _fullrank_case("disentangled", U, N, seed=31, D=D, hid=hid, hist_len=None, max_hist=40, min_hist=2, median_hist=14)
    sd["embed_distance.weight"] = sd["embed_distance.weight"] * 3  # make the distance bias matter (|coef*km| ~ 1)
    m.load_state_dict(sd)
    users = m.make_users(data.indptr, data.indices)
    got = ops.fullrank_scores("disentangled", beta, m._params(), m._catalog, users, precision=precision).cpu().numpy()
    refs = []
    for u in range(U):
        ref, scale = util.oracle_user_scores(sd, "disentangled", beta, data.coords, data.region, data.history(u), np.arange(N))
        refs.append(ref)
        assert util.cond_err(got[u], ref, scale) < util.TOL, (u, util.cond_err(got[u], ref, scale))
    s, ids = m.predict_topk(users, 10, precision=precision)
    assert ids.shape == (U, 10) and (ids >= 0).all()
    for u in range(U):
        ref = 1.0 / (1.0 + np.exp(-refs[u]))
        ref[data.history(u)] = -1.0
        util.lists_equal_outside_ties(ids[u].cpu().numpy(), s[u].cpu().numpy(), dict(enumerate(ref.tolist())), 10)


def test_fullrank_disentangled_auto_takes_the_tensor_path_and_shards():
    """precision="auto" resolves to tc_auto for the two-branch model too; two POI-range shards merge to the unsharded list."""
    U, N, beta, k = 5, 3000, 0.5, 20
    data, sd, m = _fullrank_case("disentangled", U, N, seed=33, D=64, hid=64, hist_len=None, max_hist=60, min_hist=3, median_hist=20)
    assert m.ranking_plan("auto").precision == "tc_auto"
    users = m.make_users(data.indptr, data.indices)
    s0, i0 = m.predict_topk(users, k)
    s1, i1 = m.predict_topk(users, k, precision="fp32")
    for u in range(U):
        ref, _ = util.oracle_user_scores(sd, "disentangled", beta, data.coords, data.region, data.history(u), np.arange(N))
        ref = 1.0 / (1.0 + np.exp(-ref))
        ref[data.history(u)] = -1.0
        util.lists_equal_outside_ties(i0[u].cpu().numpy(), s0[u].cpu().numpy(), dict(enumerate(ref.tolist())), k)
        util.lists_equal_outside_ties(i1[u].cpu().numpy(), s1[u].cpu().numpy(), dict(enumerate(ref.tolist())), k)
    cut = 1408  # a tile boundary is not required
    sa, ia = m.predict_topk(users, k, poi_begin=0, poi_end=cut)
    sb, ib = m.predict_topk(users, k, poi_begin=cut, poi_end=N)
    both_s, both_i = torch.cat([sa, sb], 1), torch.cat([ia, ib], 1)
    order = torch.argsort(both_s, dim=1, descending=True, stable=True)[:, :k]
    assert torch.equal(torch.gather(both_i, 1, order), i0) and torch.equal(torch.gather(both_s, 1, order), s0)


@pytest.mark.parametrize("precision", ["fp32", "tc_split", "tc_mix", "tc_auto"])
def test_fullrank_topk_matches_reference_validation_golden(precision):
    z = util.load_golden("validation_rd.npz")
    sd = util.golden_sd(z, "sd.")
    beta, U = float(z["beta"]), int(z["U"])
    m = util.make_model("region_distance", sd, beta)
    m.set_catalog(region=z["region"], coords=z["coords"])
    score, ids = m.predict_topk((z["indptr"], z["indices"]), 50, precision=precision)
    ids = ids.cpu().numpy()
    cat = orc.Catalog(z["coords"], z["region"])
    for u in range(U):
        hist = z["indices"][z["indptr"][u]:z["indptr"][u + 1]]
        _, _, cand, pred = orc.fullrank_user(sd, "region_distance", beta, cat, hist, 50, dtype=torch.float64, return_all=True)
        assert not set(ids[u].tolist()) & set(hist.tolist())  # history excluded
        util.lists_equal_outside_ties(ids[u], None, dict(zip(cand.tolist(), pred.tolist())), 50)
    # the reference's own list: identical here (no ties within tolerance in this fixture)
    assert np.array_equal(ids, z["rec"])
    # metrics through the drop-in validator == reference metrics, bit for bit
    from poi_recommendation_models_b200 import validation as V
    import types
    val = [z["val_flat"][z["val_ptr"][u]:z["val_ptr"][u + 1]].tolist() for u in range(U)]
    test = [z["test_flat"][z["test_ptr"][u]:z["test_ptr"][u + 1]].tolist() for u in range(U)]
    import scipy.sparse as sp
    csr = sp.csr_matrix((np.ones(len(z["indices"])), z["indices"], z["indptr"]), shape=(U, int(z["N"])))
    res = V.NAIS_region_distance_validation(m, types.SimpleNamespace(topk=50, powerlaw_weight=0.2), U, test, val, csr,
                                            z["region"], z["coords"], z["k_list"].tolist(), precision=precision)
    assert np.array_equal(np.array(res, dtype=np.float64), z["metrics"])


@pytest.mark.parametrize("precision", ["fp32", "tc_split", "tc_mix", "tc_fast"])
@pytest.mark.parametrize("U,N,k", [(3, 100, 50), (2, 130, 128), (300, 1000, 20), (1, 5000, 10)])
def test_fullrank_topk_shapes_and_merge(U, N, k, precision):
    beta = 0.5
    data, sd, m = _fullrank_case("region_distance", U, N, seed=U + N, hist_len=None, max_hist=min(60, N // 2), min_hist=1,
                                 median_hist=10)
    users = m.make_users(data.indptr, data.indices)
    s_all, i_all = ops.fullrank_topk("region_distance", beta, m._params(), m._catalog, users, k, precision=precision)
    scores = ops.fullrank_scores("region_distance", beta, m._params(), m._catalog, users, precision=precision).cpu().numpy()
    s_all, i_all = s_all.cpu().numpy(), i_all.cpu().numpy()
    for u in range(min(U, 8)):
        sc = scores[u].copy()
        sc[data.history(u)] = -np.inf
        order = np.lexsort((np.arange(N), -sc))  # score desc, id asc
        n_valid = N - len(data.history(u))
        kk = min(k, n_valid)
        assert np.array_equal(i_all[u, :kk], order[:kk].astype(np.int32))
        assert np.array_equal(s_all[u, :kk], sc[order[:kk]])
        assert (i_all[u, kk:] == -1).all() and np.isneginf(s_all[u, kk:]).all()
    # catalogue shards + merge == single range (the multi-GPU path, emulated on one device)
    cuts = [0, N // 3, N // 3 + 1, N]
    parts = [ops.fullrank_topk("region_distance", beta, m._params(), m._catalog, users, k, cuts[i], cuts[i + 1], precision=precision)
             for i in range(3)]
    ms, mi = ops.topk_merge(torch.stack([p[0] for p in parts], 1), torch.stack([p[1] for p in parts], 1))
    assert np.array_equal(mi.cpu().numpy(), i_all) and np.array_equal(ms.cpu().numpy(), s_all)


def test_fullrank_empty_inputs():
    data, sd, m = _fullrank_case("region_distance", 2, 200, seed=5, hist_len=4)
    users = m.make_users(data.indptr[:1], data.indices[:0])
    s, i = ops.fullrank_topk("region_distance", 0.5, m._params(), m._catalog, users, 5)
    assert s.shape == (0, 5) and i.shape == (0, 5)
    users = m.make_users(data.indptr, data.indices)
    s, i = ops.fullrank_topk("region_distance", 0.5, m._params(), m._catalog, users, 5, 10, 10)
    assert (i.cpu() == -1).all()


def test_device_metrics_bit_identical_to_reference_functions():
    from poi_recommendation_models_b200 import eval_metrics as PM
    rng = np.random.default_rng(4)
    U, k_list = 300, [5, 10, 15, 20, 25, 30]
    actual = [rng.choice(600, rng.integers(0, 7), replace=False).tolist() for _ in range(U)]
    actual[3] = actual[3] + actual[3][:1]  # duplicate positive: a Python set counts it once
    rec = np.stack([rng.choice(600, 50, replace=False) for _ in range(U)]).astype(np.int64)
    rec[5, 40:] = -1  # padded list
    got = PM.evaluate_device(actual, torch.from_numpy(rec).cuda(), k_list)
    rec_l = [[i for i in r if i >= 0] for r in rec.tolist()]
    ref = (
        [orc.precision_at_k(actual, rec_l, k) for k in k_list],
        [orc.recall_at_k(actual, rec_l, k) for k in k_list],
        [orc.hitrate_at_k(actual, rec_l, k) for k in k_list],
    )
    assert got[0] == ref[0] and got[1] == ref[1] and got[2] == ref[2]

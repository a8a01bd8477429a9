"""Property-style GPU parity cases the reference never tested: beta in {0, 0.5, 1}, duplicate history
items, H = 1, exact score ties in the top-k, saturated sigmoid, large logits, and shard/merge invariance at random cuts."""
import numpy as np
import pytest
import torch

import nais_testutil as util
from oracle import nais_oracle as orc
from poi_recommendation_models_b200 import ops, synthetic

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _case(N=500, D=64, hid=64, seed=0, style="trained", variant="region_distance"):
    coords, region, R = synthetic.make_catalog(N, seed=seed)
    sd = orc.init_state(variant, N, D, hid, R, 1, seed=seed + 1, style=style)
    return coords, region, R, sd


@pytest.mark.parametrize("beta", [0.0, 0.5, 1.0, 0.25])
def test_beta_values_pairs_and_fullrank(beta):
    N = 500
    coords, region, R, sd = _case(N)
    rng = np.random.default_rng(int(beta * 100))
    B, H = 33, 21
    hist = np.stack([rng.choice(N, H, replace=False) for _ in range(B)]).astype(np.int64)
    tgt = rng.integers(0, N, B).astype(np.int64)
    tgt[::3] = hist[::3, 0]
    aux = orc.latlon_abs_diff(coords, tgt, hist)
    m = util.make_model("region_distance", sd, beta)
    with torch.no_grad():
        s = m.attention_network(_dev(hist), _dev(tgt), _dev(region[hist]), _dev(region[tgt]), _dev(aux)).cpu().numpy()
    ref, scale = orc.attention_network_with_scale(sd, "region_distance", beta, torch.from_numpy(hist), torch.from_numpy(tgt),
                                                  torch.from_numpy(region[hist]), torch.from_numpy(region[tgt]),
                                                  torch.from_numpy(aux), dtype=torch.float64)
    assert util.cond_err(s, ref.numpy(), scale.numpy()) < util.TOL
    m.set_catalog(region=region, coords=coords)
    users = m.make_users(np.array([0, H]), hist[0])
    for prec in ("fp32", "tc_split", "tc_mix"):
        got = ops.fullrank_scores("region_distance", beta, m._params(), m._catalog, users, precision=prec).cpu().numpy()[0]
        r2, sc2 = util.oracle_user_scores(sd, "region_distance", beta, coords, region, hist[0], np.arange(N))
        assert util.cond_err(got, r2, sc2) < util.TOL, prec


def test_duplicate_history_items_and_h1():
    """The reference never pads or de-duplicates: a POI listed twice counts twice; with H = 1 the softmax has one term."""
    N = 300
    coords, region, R, sd = _case(N, seed=3)
    m = util.make_model("region_distance", sd, 0.5)
    hist = np.array([[5, 9, 5, 7, 9, 5]], dtype=np.int64).repeat(4, 0)
    tgt = np.array([1, 5, 9, 200], dtype=np.int64)  # 5 and 9 are masked at every occurrence
    aux = orc.latlon_abs_diff(coords, tgt, hist)
    with torch.no_grad():
        s = m.attention_network(_dev(hist), _dev(tgt), _dev(region[hist]), _dev(region[tgt]), _dev(aux)).cpu().numpy()
    ref, scale = orc.attention_network_with_scale(sd, "region_distance", 0.5, torch.from_numpy(hist), torch.from_numpy(tgt),
                                                  torch.from_numpy(region[hist]), torch.from_numpy(region[tgt]),
                                                  torch.from_numpy(aux), dtype=torch.float64)
    assert util.cond_err(s, ref.numpy(), scale.numpy()) < util.TOL
    m.set_catalog(region=region, coords=coords)
    for prec in ("fp32", "tc_split", "tc_auto", "tc_fast"):  # (tc_mix alone is specified for H >= 16: tc_auto routes these to SPLIT)
        users = m.make_users(np.array([0, 6, 7]), np.array([5, 9, 5, 7, 9, 5, 42]))  # user 1 has H = 1
        got = ops.fullrank_scores("region_distance", 0.5, m._params(), m._catalog, users, precision=prec).cpu().numpy()
        for u, h in enumerate(([5, 9, 5, 7, 9, 5], [42])):
            r2, sc2 = util.oracle_user_scores(sd, "region_distance", 0.5, coords, region, np.array(h), np.arange(N))
            ok = ~np.isnan(r2)  # H = 1 and target == that item: NaN in the reference too
            assert util.cond_err(got[u][ok], r2[ok], sc2[ok]) < (5e-4 if prec == "tc_fast" else util.TOL), (prec, u)
            assert np.isnan(got[u][~ok]).all()
        s_k, i_k = ops.fullrank_topk("region_distance", 0.5, m._params(), m._catalog, users, 10, precision=prec)
        assert not set(i_k[0].tolist()) & {5, 7, 9} and 42 not in i_k[1].tolist()


def test_exact_ties_are_broken_by_poi_id():
    """Identical candidate rows (same target embedding, region, coordinates) score identically; the top-k must list them
    in ascending id order on every path and across shard merges."""
    N = 400
    coords, region, R, sd = _case(N, seed=5)
    hist = np.array([10, 20, 30, 40, 50, 60, 70])
    m0 = util.make_model("region_distance", sd, 0.5)
    m0.set_catalog(region=region, coords=coords)
    sc0 = ops.fullrank_scores("region_distance", 0.5, m0._params(), m0._catalog, m0.make_users(np.array([0, len(hist)]), hist)).cpu().numpy()[0]
    sc0[hist] = -np.inf
    base = int(np.argmax(sc0))  # the user's best candidate ...
    clones = np.array([base] + [c for c in (230, 3, 399, 120) if c != base])
    for c in clones[1:]:  # ... cloned into four other ids
        sd["embed_target.weight"][c] = sd["embed_target.weight"][base]
        region[c] = region[base]
        coords[c] = coords[base]
    m = util.make_model("region_distance", sd, 0.5)
    m.set_catalog(region=region, coords=coords)
    users = m.make_users(np.array([0, len(hist)]), hist)
    for prec in ("fp32", "tc_split", "tc_mix", "tc_fast"):
        sc = ops.fullrank_scores("region_distance", 0.5, m._params(), m._catalog, users, precision=prec).cpu().numpy()[0]
        assert len(set(sc[clones].tolist())) == 1, prec  # bitwise equal scores
        s_k, i_k = ops.fullrank_topk("region_distance", 0.5, m._params(), m._catalog, users, 8, precision=prec)
        ids = i_k[0].cpu().numpy()
        assert ids[:len(clones)].tolist() == np.sort(clones).tolist(), (prec, ids)
        cuts = [0, 100, 250, N]
        parts = [ops.fullrank_topk("region_distance", 0.5, m._params(), m._catalog, users, 8, cuts[i], cuts[i + 1], precision=prec) for i in range(3)]
        ms, mi = ops.topk_merge(torch.stack([p[0] for p in parts], 1), torch.stack([p[1] for p in parts], 1))
        assert torch.equal(mi, i_k) and torch.equal(ms, s_k)


def test_saturated_sigmoid_and_large_logits():
    """forward() = sigmoid(score) saturates to exactly 1.0 / 0.0 in fp32 like the reference; ranking uses the raw score, a
    valid refinement of ranking the saturated values."""
    N = 300
    coords, region, R, sd = _case(N, seed=7)
    for k in ("embed_history.weight", "embed_target.weight", "embed_region.weight"):
        sd[k] = sd[k] * 4.0  # |score| up to hundreds, logits of tens
    rng = np.random.default_rng(1)
    B, H = 64, 9
    hist = np.stack([rng.choice(N, H, replace=False) for _ in range(B)]).astype(np.int64)
    tgt = rng.integers(0, N, B).astype(np.int64)
    aux = orc.latlon_abs_diff(coords, tgt, hist)
    m = util.make_model("region_distance", sd, 0.5)
    with torch.no_grad():
        f = m(_dev(hist), _dev(tgt), _dev(region[hist]), _dev(region[tgt]), _dev(aux)).cpu().numpy()
        s = m.attention_network(_dev(hist), _dev(tgt), _dev(region[hist]), _dev(region[tgt]), _dev(aux)).cpu().numpy()
    ref32 = orc.forward(sd, "region_distance", 0.5, torch.from_numpy(hist), torch.from_numpy(tgt), torch.from_numpy(region[hist]),
                        torch.from_numpy(region[tgt]), torch.from_numpy(aux)).numpy()
    ref, scale = orc.attention_network_with_scale(sd, "region_distance", 0.5, torch.from_numpy(hist), torch.from_numpy(tgt),
                                                  torch.from_numpy(region[hist]), torch.from_numpy(region[tgt]),
                                                  torch.from_numpy(aux), dtype=torch.float64)
    assert np.abs(ref.numpy()).max() > 20
    assert util.cond_err(s, ref.numpy(), scale.numpy()) < util.TOL
    np.testing.assert_allclose(f, ref32, rtol=util.TOL, atol=1e-30)
    assert ((f == 1.0) == (ref32 == 1.0)).all()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_shard_cuts_are_bit_identical(seed):
    rng = np.random.default_rng(seed)
    U, N, k = 7, 2000, 20
    data = synthetic.make_checkins(U, N, seed=seed + 40, hist_len=None, max_hist=70, min_hist=1, median_hist=20)
    sd = orc.init_state("region_distance", N, 64, 64, data.region_num, 1, seed=seed, style="trained")
    m = util.make_model("region_distance", sd, 0.5)
    m.set_catalog(region=data.region, coords=data.coords)
    users = m.make_users(data.indptr, data.indices)
    for prec in ("fp32", "tc_split", "tc_mix"):
        full = ops.fullrank_topk("region_distance", 0.5, m._params(), m._catalog, users, k, precision=prec)
        cuts = [0] + sorted(rng.choice(np.arange(1, N), 4, replace=False).tolist()) + [N]
        parts = [ops.fullrank_topk("region_distance", 0.5, m._params(), m._catalog, users, k, cuts[i], cuts[i + 1], precision=prec)
                 for i in range(len(cuts) - 1)]
        ms, mi = ops.topk_merge(torch.stack([p[0] for p in parts], 1), torch.stack([p[1] for p in parts], 1))
        assert torch.equal(mi, full[1]) and torch.equal(ms, full[0]), (prec, cuts)


@pytest.mark.parametrize("H", [513, 700])
def test_long_history_beyond_the_smem_staging_window(H):
    """History longer than the 512 items the tensor kernel stages in shared memory (the rest is read with __ldg), odd
    length (last chunk half empty)."""
    N = 1200
    coords, region, R, sd = _case(N, seed=9)
    rng = np.random.default_rng(H)
    hist = np.sort(rng.choice(N, H, replace=False))
    m = util.make_model("region_distance", sd, 0.5)
    m.set_catalog(region=region, coords=coords)
    users = m.make_users(np.array([0, H]), hist)
    ref, scale = util.oracle_user_scores(sd, "region_distance", 0.5, coords, region, hist, np.arange(N))
    for prec in ("fp32", "tc_split", "tc_mix"):
        got = ops.fullrank_scores("region_distance", 0.5, m._params(), m._catalog, users, precision=prec).cpu().numpy()[0]
        assert util.cond_err(got, ref, scale) < util.TOL, prec


@pytest.mark.parametrize("D,hid", [(64, 128), (64, 96), (32, 128), (48, 32), (16, 16), (32, 64), (128, 128), (128, 64), (96, 96),
                                   (128, 32), (96, 16), (256, 256), (256, 128), (128, 256), (256, 64), (64, 256), (192, 96)])
@pytest.mark.parametrize("precision", ["tc_split", "tc_mix", "tc_fast"])
def test_tensor_path_shapes(D, hid, precision):
    """Every (D, hid) tiling of the tensor-core path: two history items per MMA step for hid <= 64, one for hid 96/128
    (a cell's hidden columns are split between two warps and exchanged through shared memory); hid = 256 as two 128-unit halves
    in consecutive chunks (the two epilogue groups meet once per cell); D > 128 as up to eight K-parts with one resident
    candidate tile."""
    U, N = 4, 900
    if precision == "tc_mix" and D % 32:
        pytest.skip("tc_mix: an e5m2 MMA covers K = 32")
    data = synthetic.make_checkins(U, N, seed=D + hid, hist_len=None, max_hist=50, min_hist=1, median_hist=15)
    sd = orc.init_state("region_distance", N, D, hid, data.region_num, 1, seed=3, style="trained")
    m = util.make_model("region_distance", sd, 0.5)
    m.set_catalog(region=data.region, coords=data.coords)
    users = m.make_users(data.indptr, data.indices)
    got = ops.fullrank_scores("region_distance", 0.5, m._params(), m._catalog, users, precision=precision).cpu().numpy()
    tol = 5e-4 if precision == "tc_fast" else util.TOL
    for u in range(U):
        ref, scale = util.oracle_user_scores(sd, "region_distance", 0.5, data.coords, data.region, data.history(u), np.arange(N))
        assert util.cond_err(got[u], ref, scale) < tol, (u, util.cond_err(got[u], ref, scale))
    full = ops.fullrank_topk("region_distance", 0.5, m._params(), m._catalog, users, 20, precision=precision)
    parts = [ops.fullrank_topk("region_distance", 0.5, m._params(), m._catalog, users, 20, lo, hi, precision=precision)
             for lo, hi in ((0, 300), (300, 333), (333, N))]
    ms, mi = ops.topk_merge(torch.stack([p[0] for p in parts], 1), torch.stack([p[1] for p in parts], 1))
    assert torch.equal(mi, full[1]) and torch.equal(ms, full[0])


@pytest.mark.parametrize("D,hid", [(256, 64), (256, 256), (128, 128), (192, 96)])
def test_fp32_fullrank_large_dims(D, hid):
    """embed_size up to 256 through the fused FP32 full-rank kernel (above 128 the candidate tile is not staged in smem)."""
    U, N = 3, 500
    data = synthetic.make_checkins(U, N, seed=D, hist_len=None, max_hist=40, min_hist=1, median_hist=12)
    sd = orc.init_state("region_distance", N, D, hid, data.region_num, 1, seed=3, style="trained")
    m = util.make_model("region_distance", sd, 0.5)
    m.set_catalog(region=data.region, coords=data.coords)
    users = m.make_users(data.indptr, data.indices)
    got = ops.fullrank_scores("region_distance", 0.5, m._params(), m._catalog, users, precision="fp32").cpu().numpy()
    for u in range(U):
        ref, scale = util.oracle_user_scores(sd, "region_distance", 0.5, data.coords, data.region, data.history(u), np.arange(N))
        assert util.cond_err(got[u], ref, scale) < util.TOL, (u, util.cond_err(got[u], ref, scale))


def test_predict_topk_slices_huge_batches(monkeypatch):
    """Batches whose workspace would exceed ops.WORKSPACE_LIMIT_BYTES are scored in user slices with identical results."""
    U, N = 40, 800
    data = synthetic.make_checkins(U, N, seed=77, hist_len=None, max_hist=40, min_hist=1, median_hist=12)
    sd = orc.init_state("region_distance", N, 64, 64, data.region_num, 1, seed=3, style="trained")
    m = util.make_model("region_distance", sd, 0.5)
    m.set_catalog(region=data.region, coords=data.coords)
    users = m.make_users(data.indptr, data.indices)
    full = ops.fullrank_topk("region_distance", 0.5, m._params(), m._catalog, users, 10, precision="tc_split")
    monkeypatch.setattr(ops, "WORKSPACE_LIMIT_BYTES", 8 << 20)
    sliced = ops.fullrank_topk("region_distance", 0.5, m._params(), m._catalog, users, 10, precision="tc_split")
    assert torch.equal(full[0], sliced[0]) and torch.equal(full[1], sliced[1])


def test_tc_auto_gate_per_user_and_weight_scale():
    """precision="tc_auto": a user takes the fp16 + e5m2-correction kernels (MIX) when the device-side bound
    rho = max|p| * max|B| * sqrt(hid * D) is <= 128 AND the history has >= 16 items (MIX's per-term error averages out over
    the history), else the three-pass fp16 split.  Each user's row is bit-identical to calling that mode directly, and
    within tolerance of the float64 oracle; beyond the bound everything falls back to SPLIT."""
    N = 600
    lens = [3, 40, 15, 16, 2, 33]  # (H = 1 is covered by test_duplicate_history_items_and_h1; there |S| itself cancels over d)
    U = len(lens)
    rng = np.random.default_rng(5)
    indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    indices = np.concatenate([rng.choice(N, n, replace=False) for n in lens]).astype(np.int64)
    coords, region, R = synthetic.make_catalog(N, seed=21)
    sd = orc.init_state("region_distance", N, 64, 64, R, 1, seed=4, style="trained")

    def run(sd_):
        m = util.make_model("region_distance", sd_, 0.5)
        m.set_catalog(region=region, coords=coords)
        users = m.make_users(indptr, indices)
        out = {p: ops.fullrank_scores("region_distance", 0.5, m._params(), m._catalog, users, precision=p).cpu().numpy()
               for p in ("tc_split", "tc_mix")}
        s_auto, i_auto = ops.fullrank_topk("region_distance", 0.5, m._params(), m._catalog, users, 10, precision="tc_auto")
        choice = ops.last_tc_choice()
        out["tc_auto"] = ops.fullrank_scores("region_distance", 0.5, m._params(), m._catalog, users, precision="tc_auto").cpu().numpy()
        tk = {p: ops.fullrank_topk("region_distance", 0.5, m._params(), m._catalog, users, 10, precision=p) for p in ("tc_split", "tc_mix")}
        err = 0.0
        for u in range(U):
            same = "tc_mix" if choice["use_mix"] and lens[u] >= choice["min_hist_for_mix"] else "tc_split"
            assert np.array_equal(out["tc_auto"][u], out[same][u], equal_nan=True), (u, same)
            assert torch.equal(i_auto[u], tk[same][1][u]) and torch.equal(s_auto[u], tk[same][0][u]), (u, same)
            hist = indices[indptr[u]:indptr[u + 1]]
            ref, scale = util.oracle_user_scores(sd_, "region_distance", 0.5, coords, region, hist, np.arange(N))
            ok = ~np.isnan(ref)
            err = max(err, util.cond_err(out["tc_auto"][u][ok], ref[ok], scale[ok]))
        return choice, err

    choice, err = run(sd)
    assert choice["use_mix"] == 1 and choice["rho"] < 128 and choice["min_hist_for_mix"] == 16, choice
    assert err < util.TOL, err
    big = {k: v.clone() for k, v in sd.items()}
    for k in big:
        if k.startswith("embed_"):
            big[k] = big[k] * (1.0 / 0.3)          # embedding std 1.0
    big["attn_layer1.weight"] = big["attn_layer1.weight"] * 5
    big["attn_layer2.weight"] = big["attn_layer2.weight"] * 5
    choice, err = run(big)
    assert choice["use_mix"] == 0 and choice["rho"] > 256, choice
    assert err < util.TOL, err


def test_fullrank_topk_is_cuda_graph_capturable():
    """The C ABI only enqueues on the given stream (no allocation, no synchronisation, no host read-back — the tc_auto
    precision gate runs on the device), so a whole ranking call can be captured in a CUDA graph and replayed."""
    N = 1500
    coords, region, R, sd = _case(N, seed=9)
    m = util.make_model("region_distance", sd, 0.5)
    m.set_catalog(region=region, coords=coords)
    rng = np.random.default_rng(1)
    lens = [20, 5, 33, 16]
    indptr = np.concatenate([[0], np.cumsum(lens)])
    users = m.make_users(indptr, np.concatenate([rng.choice(N, n, replace=False) for n in lens]))
    P = m._params()
    for prec in ("tc_auto", "fp32"):
        eager = ops.fullrank_topk("region_distance", 0.5, P, m._catalog, users, 10, precision=prec)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            ops.fullrank_topk("region_distance", 0.5, P, m._catalog, users, 10, precision=prec)  # warm-up outside the capture
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            s_g, i_g = ops.fullrank_topk("region_distance", 0.5, P, m._catalog, users, 10, precision=prec)
        s_g.fill_(0)
        i_g.fill_(0)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(i_g, eager[1]) and torch.equal(s_g, eager[0]), prec


@pytest.mark.parametrize("prec", ["fp32", "tc_split", "tc_auto"])
def test_more_than_65535_users_in_one_call(prec):
    """User batches beyond CUDA's 65 535 grid.y limit: users ride grid.x in every kernel of the ranking call.  66 000 short
    users against a thin catalogue; the lists of the first / middle / last users equal a small call on those users alone."""
    N, U, H, k = 256, 66000, 3, 5
    coords, region, R, sd = _case(N, seed=11)
    m = util.make_model("region_distance", sd, 0.5)
    m.set_catalog(region=region, coords=coords)
    rng = np.random.default_rng(12)
    hist = (rng.integers(0, N - H, (U, 1)) + np.arange(H)[None, :]).astype(np.int64)  # distinct items per user
    indptr = np.arange(0, (U + 1) * H, H, dtype=np.int64)
    s_all, i_all = m.predict_topk((indptr, hist.reshape(-1)), k, precision=prec)
    pick = np.array([0, 1, 32767, 65534, 65535, 65536, U - 1])
    s_few, i_few = m.predict_topk((np.arange(0, (len(pick) + 1) * H, H, dtype=np.int64), hist[pick].reshape(-1)), k, precision=prec)
    torch.cuda.synchronize()
    assert torch.equal(i_all[pick], i_few)
    assert torch.equal(s_all[pick], s_few)
    from poi_recommendation_models_b200 import powerlaw
    geo = powerlaw.log_scores(m._catalog, m.make_users(indptr, hist.reshape(-1)), 0.5, -1.2)
    few = powerlaw.log_scores(m._catalog, m.make_users(np.arange(0, (len(pick) + 1) * H, H, dtype=np.int64), hist[pick].reshape(-1)), 0.5, -1.2)
    assert torch.equal(geo[pick], few)

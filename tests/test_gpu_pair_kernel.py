"""The CTA-pair (tcgen05 cta_group::2) full-rank kernels (csrc/nais_tc.cu kFix 5 / 6; the default for D = hid = 64, opt-out
NAIS_PREC_FLAG_ONE_CTA = precision names ending in "_onecta"): two CTAs of a cluster run one M = 256 MMA per step, each staging
half of every user-operand chunk.  Same products, same summation order per
accumulator element as the one-CTA kernels: scores and lists must be IDENTICAL, for even and odd numbers of candidate-tile
groups, ragged histories (incl. the MIX / SPLIT split of tc_auto), a catalogue range that ends inside a tile, and POI-range
shards."""
import numpy as np
import pytest
import torch

import nais_testutil as util
from oracle import nais_oracle as orc
from poi_recommendation_models_b200 import ops, synthetic

pytestmark = pytest.mark.gpu


def _case(N, U, seed, max_hist=140):
    data = synthetic.make_checkins(U, N, seed=seed, hist_len=None, max_hist=max_hist, min_hist=2, median_hist=30)
    sd = orc.init_state("region_distance", N, 64, 64, data.region_num, 1, seed=seed + 1, style="trained")
    m = util.make_model("region_distance", sd, 0.5)
    m.set_catalog(region=data.region, coords=data.coords)
    return data, sd, m


@pytest.mark.parametrize("precision", ["tc_mix", "tc_split", "tc_auto"])
@pytest.mark.parametrize("N", [384 * 2, 384 * 5 + 100, 384 * 6 + 1, 40000])
def test_pair_kernel_is_bit_identical_to_the_one_cta_kernel(precision, N):
    U = 9 if N > 10000 else 23
    data, sd, m = _case(N, U, seed=N % 1000)
    users = m.make_users(data.indptr, data.indices)
    s0 = ops.fullrank_scores("region_distance", 0.5, m._params(), m._catalog, users, precision=precision + "_onecta")
    s1 = ops.fullrank_scores("region_distance", 0.5, m._params(), m._catalog, users, precision=precision)
    assert torch.equal(s0, s1), float((s0 - s1).abs().max())
    a = m.predict_topk(users, 20, precision=precision + "_onecta")
    b = m.predict_topk(users, 20, precision=precision)
    assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0])
    k50 = m.predict_topk(users, 50, precision=precision)  # k > 32: the shared-memory sort path
    k50_ref = m.predict_topk(users, 50, precision=precision + "_onecta")
    assert torch.equal(k50[1], k50_ref[1]) and torch.equal(k50[0], k50_ref[0])


def test_pair_kernel_against_the_oracle_and_on_shards():
    N, U, k = 5000, 12, 20
    data, sd, m = _case(N, U, seed=77)
    users = m.make_users(data.indptr, data.indices)
    got = ops.fullrank_scores("region_distance", 0.5, m._params(), m._catalog, users, precision="tc_auto").cpu().numpy()
    for u in range(0, U, 3):
        ref, scale = util.oracle_user_scores(sd, "region_distance", 0.5, data.coords, data.region, data.history(u), np.arange(N))
        assert util.cond_err(got[u], ref, scale) < util.TOL
    whole = m.predict_topk(users, k, precision="tc_auto")
    cut = 2000
    sa, ia = m.predict_topk(users, k, poi_begin=0, poi_end=cut, precision="tc_auto")
    sb, ib = m.predict_topk(users, k, poi_begin=cut, poi_end=N, precision="tc_auto")
    both_s, both_i = torch.cat([sa, sb], 1), torch.cat([ia, ib], 1)
    order = torch.argsort(both_s, dim=1, descending=True, stable=True)[:, :k]
    assert torch.equal(torch.gather(both_i, 1, order), whole[1]) and torch.equal(torch.gather(both_s, 1, order), whole[0])


def test_one_cta_flag_is_ignored_where_there_is_no_pair_kernel():
    N, U = 900, 4
    data = synthetic.make_checkins(U, N, seed=5, hist_len=None, max_hist=30, min_hist=2, median_hist=10)
    sd = orc.init_state("region_distance", N, 32, 32, data.region_num, 1, seed=6, style="trained")
    m = util.make_model("region_distance", sd, 0.5)
    m.set_catalog(region=data.region, coords=data.coords)
    users = m.make_users(data.indptr, data.indices)
    a = m.predict_topk(users, 10, precision="tc_auto")
    b = m.predict_topk(users, 10, precision="tc_auto_onecta")
    assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0])


def test_cta_pair_building_blocks_against_a_cpu_gemm():
    """tests/umma_probe_pair.cu (built by __graft_entry__.build()): tcgen05 alloc / mma / commit with cta_group::2 on a cluster of
    two CTAs — M = 256, N = 144 with each CTA staging half of B, the relay barrier, a generic-written A chunk, an e5m2 pass, the
    multicast commit — every accumulator element equal to a CPU GEMM."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "umma_probe_pair.bin")
    if not os.path.isfile(exe):
        pytest.skip("probe binary not built (run __graft_entry__.build())")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "PAIR PROBE OK" in r.stdout, r.stdout + r.stderr

"""Power-law geographical weighting (poi_recommendation_models_b200/powerlaw.py) against
tests/golden/powerlaw.npz — the fitted (a, b), the distance distribution and PowerLaw.predict values the UNMODIFIED
reference powerLaw.py produced (tests/golden/make_golden_powerlaw.py) — and the device log-score / re-ranking path
against the float64 restatement of that same class."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import nais_testutil as util
from oracle import nais_oracle as orc
from poi_recommendation_models_b200 import powerlaw as PL


def _load():
    z = util.load_golden("powerlaw.npz")
    U, N = int(z["U"]), int(z["N"])
    csr = sp.csr_matrix((np.ones(len(z["indices"])), z["indices"], z["indptr"]), shape=(U, N))
    return z, U, N, csr


def test_fit_and_predict_match_reference():
    z, U, N, csr = _load()
    x, t = PL.PowerLaw.compute_distance_distribution(csr, z["coords"])
    assert np.array_equal(np.asarray(x), z["x"]) and np.allclose(t, z["t"], rtol=1e-15, atol=0)
    np.random.seed(11)
    G = PL.PowerLaw()
    G.fit_distance_distribution(csr, z["coords"])
    assert abs(G.a - float(z["a"])) <= 1e-12 * abs(float(z["a"])) and abs(G.b - float(z["b"])) <= 1e-12
    pred = np.array([[G.predict(u, int(j)) for j in z["cand"]] for u in range(U)])
    np.testing.assert_allclose(pred, z["pred"], rtol=1e-10)
    # the scalar distance itself, incl. the short-circuit
    assert PL.dist((40.0, -74.0), (40.0 + 5e-7, -74.0 - 5e-7)) == 0.0
    assert abs(float(PL.dist((40.7, -74.0), (40.8, -73.9))) - float(orc.dist_km(np.array([40.7]), np.array([-74.0]), np.array([40.8]), np.array([-73.9]))[0])) < 1e-9


@pytest.mark.gpu
def test_device_log_scores_and_rerank():
    from poi_recommendation_models_b200 import ops
    z, U, N, csr = _load()
    a, b, alpha, k = 0.8, -1.3, 0.2, 15  # a steeper law than the golden fit, so the geographic term matters
    sd = orc.init_state("region_distance", N, 64, 64, int(z["region"].max()) + 1, 1, seed=8, style="trained")
    m = util.make_model("region_distance", sd, 0.5)
    m.set_catalog(region=z["region"], coords=z["coords"])
    users = m.make_users(z["indptr"], z["indices"])
    logg = PL.log_scores(m._catalog, users, a, b).cpu().numpy().astype(np.float64)
    G = PL.PowerLaw(a, b)
    G.poi_coos = z["coords"]
    G.visited_lids = {u: z["indices"][z["indptr"][u]:z["indptr"][u + 1]] for u in range(U)}
    ref = np.array([[np.log(G.predict(u, j)) for j in range(N)] for u in range(U)])
    # fp32 distances from centred coordinates: a few 1e-6 relative per factor -> 1e-4 absolute over <= 30 log terms
    np.testing.assert_allclose(logg, ref, rtol=2e-5, atol=2e-4)
    val, idx = PL.rerank_topk(m, users, k, a, b, alpha, precision="fp32")
    scores = ops.fullrank_scores("region_distance", 0.5, m._params(), m._catalog, users, precision="fp32").cpu().numpy().astype(np.float64)
    for u in range(U):
        hist = z["indices"][z["indptr"][u]:z["indptr"][u + 1]]
        cand = np.setdiff1d(np.arange(N), hist)
        g = PL.normalize(np.exp(ref[u, cand]))                     # run.py:55-59 over the candidate list
        mixed = (1 - alpha) / (1 + np.exp(-scores[u, cand])) + alpha * g
        order = cand[np.argsort(-mixed, kind="stable")[:k]]
        got = idx[u].cpu().numpy()
        assert not set(got.tolist()) & set(hist.tolist())
        ref_val = dict(zip(cand.tolist(), mixed.tolist()))
        # same list outside 1e-4 tie bands of the mixed score
        for r in range(k):
            assert abs(ref_val[int(got[r])] - ref_val[int(order[r])]) < 1e-4, (u, r)
        np.testing.assert_allclose(val[u].cpu().numpy(), [ref_val[int(j)] for j in got], rtol=1e-4, atol=1e-6)

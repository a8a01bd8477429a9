"""GPU: the sibling validators `NAIS_validation` / `NAIS_region_validation` (drop-ins for validation.py:7-59) against goldens
written by the unmodified reference validators (tests/golden/make_golden_validators.py).  (File name sorts last on purpose.)"""
import types

import numpy as np
import pytest
import scipy.sparse as sp
import torch

import nais_testutil as util
from oracle import nais_oracle as orc
from poi_recommendation_models_b200 import validation as V

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp32", "auto"])
@pytest.mark.parametrize("variant", ["basic", "region"])
def test_sibling_validators_match_reference_golden(variant, precision):
    z = util.load_golden(f"validation_{variant}.npz")
    sd = util.golden_sd(z, "sd.")
    beta, U, N = float(z["beta"]), int(z["U"]), int(z["N"])
    m = util.make_model(variant, sd, beta)
    val = [z["val_flat"][z["val_ptr"][u]:z["val_ptr"][u + 1]].tolist() for u in range(U)]
    test = [z["test_flat"][z["test_ptr"][u]:z["test_ptr"][u + 1]].tolist() for u in range(U)]
    csr = sp.csr_matrix((np.ones(len(z["indices"])), z["indices"], z["indptr"]), shape=(U, N))
    args, k_list = types.SimpleNamespace(topk=50), z["k_list"].tolist()
    if variant == "basic":
        res = V.NAIS_validation(m, args, U, test, val, csr, k_list, precision=precision)
    else:
        res = V.NAIS_region_validation(m, args, U, test, val, csr, z["region"], k_list, precision=precision)
    _, ids = m.predict_topk((z["indptr"], z["indices"]), 50, precision=precision)
    ids = ids.cpu().numpy()
    cat = orc.Catalog(z["coords"], z["region"])
    for u in range(U):
        hist = z["indices"][z["indptr"][u]:z["indptr"][u + 1]]
        _, _, cand, pred = orc.fullrank_user(sd, variant, beta, cat, hist, 50, dtype=torch.float64, return_all=True)
        assert not set(ids[u].tolist()) & set(hist.tolist())  # history excluded
        util.lists_equal_outside_ties(ids[u], None, dict(zip(cand.tolist(), pred.tolist())), 50)
    if np.array_equal(ids, z["rec"]):  # no near-tie was ordered differently: the metrics must then be the reference's, bit for bit
        assert np.array_equal(np.array(res, dtype=np.float64), z["metrics"])
    else:
        np.testing.assert_allclose(np.array(res, dtype=np.float64), z["metrics"], atol=0.05)

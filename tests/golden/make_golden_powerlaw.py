"""Generate tests/golden/powerlaw.npz by RUNNING THE UNMODIFIED REFERENCE `powerLaw.py` (PowerLaw.fit_distance_distribution,
pr_d, predict) on a small synthetic check-in matrix.      python tests/golden/make_golden_powerlaw.py   # needs /root/reference"""
from __future__ import annotations

import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from poi_recommendation_models_b200 import synthetic  # noqa: E402


def main():
    ref = ref_shim.load_reference("powerLaw")
    U, N = 12, 300
    data = synthetic.make_checkins(U, N, seed=33, hist_len=None, max_hist=30, min_hist=4, median_hist=12)
    csr = sp.csr_matrix((np.ones(len(data.indices)), data.indices, data.indptr), shape=(U, N))
    coords = np.asarray(data.coords, dtype=np.float64)
    np.random.seed(11)
    G = ref.PowerLaw()
    G.fit_distance_distribution(csr, coords)
    x, t = ref.PowerLaw.compute_distance_distribution(csr, coords)
    cand = np.arange(0, N, 7)
    pred = np.array([[G.predict(u, int(j)) for j in cand] for u in range(U)], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "powerlaw.npz"), indptr=data.indptr, indices=data.indices, coords=coords, U=U, N=N,
                        a=np.float64(G.a), b=np.float64(G.b), x=np.array(x), t=np.array(t), cand=cand, pred=pred,
                        region=data.region)
    print("a, b =", G.a, G.b, "pred range", pred.min(), pred.max())


if __name__ == "__main__":
    main()

"""Generate tests/golden/rank_metrics.npz by RUNNING THE UNMODIFIED REFERENCE `eval_metrics.py`
(precision_at_k_per_sample, apk, mapk: eval_metrics.py:29-34, 70-125) on seeded lists that include the edge cases those
functions branch on: users without positives, recommendation lists with repeated ids, lists shorter than k.
    python tests/golden/make_golden_rank_metrics.py   # needs /root/reference"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

KS = (1, 5, 10, 20, 50)


def make_lists(seed=7, users=64, pois=120, width=30):
    """Returns (actual, predicted) as lists of int lists; rows 0..3 are the hand-made edge cases."""
    rng = np.random.default_rng(seed)
    actual = [rng.choice(pois, int(rng.integers(0, 9)), replace=False).tolist() for _ in range(users)]
    predicted = [rng.choice(pois, width, replace=False).tolist() for _ in range(users)]
    actual[0], predicted[0] = [], predicted[0]                       # no positives
    actual[1], predicted[1] = [3, 4, 5], [3, 3, 4, 3, 5, 5] + predicted[1][:10]  # repeats in the ranked list
    actual[2], predicted[2] = [7, 8], [8, 1, 7]                      # list shorter than most k
    actual[3], predicted[3] = list(range(40)), list(range(30))       # more positives than k, all hits
    return actual, predicted


def main():
    ref = ref_shim.load_reference("eval_metrics")
    actual, predicted = make_lists()
    apk = np.array([[ref.apk(a, p, k) for k in KS] for a, p in zip(actual, predicted)], dtype=np.float64)
    mapk = np.array([ref.mapk(actual, predicted, k) for k in KS], dtype=np.float64)
    pps = np.array([[ref.precision_at_k_per_sample(a, p, k) for k in KS] for a, p in zip(actual, predicted)], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "rank_metrics.npz"), ks=np.array(KS), apk=apk, mapk=mapk, pps=pps)
    print("mapk", mapk)


if __name__ == "__main__":
    main()

"""Generate tests/golden/validation_basic.npz and validation_region.npz by RUNNING THE UNMODIFIED REFERENCE validators
`validation.NAIS_validation` (validation.py:7-31) and `validation.NAIS_region_validation` (validation.py:34-59) with the
reference's own `model.NAIS_basic` / `model.NAIS_regionEmbedding` and `batches.get_NAIS_batch_test*` on a tiny dataset.
    python tests/golden/make_golden_validators.py      # needs /root/reference"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import nais_oracle as orc  # noqa: E402
from oracle import ref_shim  # noqa: E402
from poi_recommendation_models_b200 import synthetic  # noqa: E402
from make_golden import _ref_model  # noqa: E402


def main():
    ref_model, ref_validation, ref_metrics = ref_shim.load_reference("model", "validation", "eval_metrics")
    torch.set_num_threads(4)
    k_list = [5, 10, 15, 20, 25, 30]
    for variant, seed in (("basic", 31), ("region", 37)):
        U, N, D, hid, beta = 5, 400, 64, 64, 0.5
        data = synthetic.make_checkins(U, N, hist_len=None, seed=seed, max_hist=35, min_hist=4, median_hist=12)
        sd = orc.init_state(variant, N, D, hid, data.region_num, 1, seed=seed + 1, style="trained")
        m = _ref_model(ref_model, variant, N, D, hid, beta, data.region_num, sd)
        captured = {}

        def capture(positive, recommended, ks):
            captured.setdefault("rec", [list(r) for r in recommended])
            return ([ref_metrics.precision_at_k(positive, recommended, k) for k in ks],
                    [ref_metrics.recall_at_k(positive, recommended, k) for k in ks],
                    [ref_metrics.hitrate_at_k(positive, recommended, k) for k in ks])

        ref_validation.eval_metrics = types.SimpleNamespace(evaluate_mp=capture)
        args = types.SimpleNamespace(topk=50)
        with torch.no_grad():
            if variant == "basic":
                res = ref_validation.NAIS_validation(m, args, U, data.test_positive, data.val_positive, data.train_csr(), k_list)
            else:
                res = ref_validation.NAIS_region_validation(m, args, U, data.test_positive, data.val_positive, data.train_csr(),
                                                            data.region, k_list)
        ref_validation.eval_metrics = ref_metrics
        out = {"U": U, "N": N, "D": D, "hid": hid, "beta": beta, "k_list": np.array(k_list),
               "rec": np.array(captured["rec"], dtype=np.int64), "metrics": np.array(res, dtype=np.float64),
               "coords": data.coords, "region": data.region, "indptr": data.indptr, "indices": data.indices,
               "val_flat": np.concatenate([np.array(v, dtype=np.int64) for v in data.val_positive]),
               "val_ptr": np.cumsum([0] + [len(v) for v in data.val_positive]),
               "test_flat": np.concatenate([np.array(v, dtype=np.int64) for v in data.test_positive]),
               "test_ptr": np.cumsum([0] + [len(v) for v in data.test_positive])}
        for k, v in sd.items():
            out["sd." + k] = v.numpy()
        np.savez_compressed(os.path.join(HERE, f"validation_{variant}.npz"), **out)
        print("wrote", variant, "recall@10 val/test:", res[1][1], res[4][1])


if __name__ == "__main__":
    main()

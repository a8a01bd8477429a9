"""Diagnostic: ELEMENT-WISE relative error of the pair backward (tcgen05 vs FP32 kernels) against the float64 oracle on a
reference-style batch (one user, history repeated per row) at the reference's init scale (embeddings N(0, 0.01)): what a
per-element optimizer (Adagrad from a zero accumulator) sees, as opposed to the per-tensor 2e-4-of-max bar of the tests."""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import nais_oracle as orc  # noqa: E402
from poi_recommendation_models_b200 import batches as PB, model as M, synthetic  # noqa: E402

U, N, D, hid = 8, 2500, 64, 64
data = synthetic.make_checkins(U, N, seed=2, hist_len=None, max_hist=60, min_hist=4, median_hist=18)
csr = data.train_csr()
for style in ("reference", "trained"):
    sd = orc.init_state("region_distance", N, D, hid, data.region_num, 1, seed=0, style=style)
    for u in (0, 3):
        random.seed(u)
        hist, tgt, label, hreg, treg = PB.get_NAIS_batch_region(csr, N, u, 4, data.region)
        ll = PB.lat_lon_pairs(data.coords, tgt.cpu().numpy(), hist[0].cpu().numpy())
        ref_in = [t.cpu() for t in (hist, tgt, hreg, treg, ll)]
        pred = torch.sigmoid(orc.attention_network(sd, "region_distance", 0.5, *ref_in, dtype=torch.float64))
        dscore = ((pred - label.cpu().double()) / len(label)).numpy()  # d mean-BCE / d score
        _, ref = orc.grads(sd, "region_distance", 0.5, *ref_in, torch.from_numpy(dscore))
        for pp in ("tc", "fp32"):
            m = M.NAIS_region_distance_Embedding(N, D, hid, 0.5, data.region_num, 1)
            m.load_state_dict(sd)
            m = m.cuda()
            m.pairs_precision = pp
            s = m.attention_network(hist, tgt, hreg, treg, ll)
            (s * torch.from_numpy(dscore).float().cuda()).sum().backward()
            out = {}
            for name, p in m.named_parameters():
                if name not in ref or p.grad is None:
                    continue
                r = ref[name].numpy().astype(np.float64)
                g = p.grad.detach().cpu().double().numpy()
                nz = np.abs(r) > 0
                rel = np.abs(g - r)[nz] / np.abs(r)[nz]
                flips = float(np.mean(np.sign(g[nz]) != np.sign(r[nz])))
                out[name.replace(".weight", "")] = f"med {np.median(rel):.1e} p99 {np.quantile(rel, 0.99):.1e} flips {flips:.1e}"
            print(style, "user", u, "H", hist.shape[1], pp, out, flush=True)

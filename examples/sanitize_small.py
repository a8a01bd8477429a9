#!/usr/bin/env python
"""One small launch of every kernel in libnais_b200.so — the workload for `compute-sanitizer --tool memcheck|racecheck`."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poi_recommendation_models_b200 import model as M, ops, synthetic


def main():
    dev = torch.device("cuda:0")
    N, D, hid = 700, 64, 64
    data = synthetic.make_checkins(6, N, seed=3, hist_len=None, max_hist=40, min_hist=2, median_hist=12)
    reg = torch.from_numpy(data.region).to(dev)
    c = torch.from_numpy(data.coords).to(dev)
    rng = np.random.default_rng(0)
    for name in ("NAIS_region_distance_Embedding", "NAIS_regionEmbedding", "NAIS_region_distance_disentangled_Embedding"):
        torch.manual_seed(0)
        cls = getattr(M, name)
        m = (cls(N, D, hid, 0.5, data.region_num) if name == "NAIS_regionEmbedding" else cls(N, D, hid, 0.5, data.region_num, 1)).to(dev)
        with torch.no_grad():
            for n_, p in m.named_parameters():
                if n_.startswith("embed_"):
                    p.normal_(0, 0.3)
        B, H = 37, 13
        hist = torch.from_numpy(np.stack([rng.choice(N, H, replace=False) for _ in range(B)])).to(dev)
        tgt = torch.from_numpy(rng.integers(0, N, B)).to(dev)
        ll = (c[tgt][:, None, :] - c[hist]).abs().float().contiguous()
        m.train()
        if name == "NAIS_regionEmbedding":
            s = m.attention_network(hist, tgt, reg[hist], reg[tgt])  # train-mode dropout path
        elif "disentangled" in name:
            s = m.attention_network(hist, tgt, reg[hist], reg[tgt], ll[..., 0].contiguous() * 100)
        else:
            s = m.attention_network(hist, tgt, reg[hist], reg[tgt], ll)
        s.sum().backward()
        if name == "NAIS_region_distance_Embedding":  # the run above took the tensor-core pair kernels (default); now the FP32 ones
            m.pairs_precision = "fp32"
            m.zero_grad(set_to_none=True)
            m.attention_network(hist, tgt, reg[hist], reg[tgt], ll).sum().backward()
            m.pairs_precision = "auto"
        m.eval()
        m.set_catalog(region=data.region, coords=data.coords)
        users = m.make_users(data.indptr, data.indices)
        # (two branches: one tensor pass per branch + haversine in the epilogue; one branch: the CTA-pair and the one-CTA kernels)
        precs = ("fp32", "tc_split", "tc_mix") if "disentangled" in name else ("fp32", "tc_split", "tc_mix", "tc_fast", "tc_auto_onecta")
        outs = []
        for prec in precs:
            outs.append(ops.fullrank_topk(m.variant, 0.5, m._params(), m._catalog, users, 10, precision=prec))
            ops.fullrank_scores(m.variant, 0.5, m._params(), m._catalog, users, 0, 300, precision=prec)
        ops.topk_merge(torch.stack([o[0] for o in outs], 1), torch.stack([o[1] for o in outs], 1))
        if "disentangled" not in name:  # the one-call optimizer step and the in-library user loop (device sampler, BCE, Adagrad)
            from poi_recommendation_models_b200 import batches as PB
            m.train()
            if hasattr(m, "drop"):
                m.drop.p = 0.0
            opt = torch.optim.Adagrad(m.parameters(), lr=0.01)
            bt = PB.DeviceBatcher(data.train_csr(), data.region, data.coords, device=dev, seed=0)
            b = bt.multi_user_batch(np.arange(4), 4, seed=1)
            m.fused_adagrad_step(opt, b.label, b)
            m.train_users(opt, bt, np.arange(6), 4, seed=2)
        ops.check_indices(sync=True)
        torch.cuda.synchronize()
        print(name, "ok", float(s.detach().sum()))


if __name__ == "__main__":
    main()

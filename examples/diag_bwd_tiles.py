"""Diagnostic: tensor-core pair backward vs the FP32 backward as the number of rows crosses the persistent grid (296 CTAs)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nais_testutil as util  # noqa: E402
from oracle import nais_oracle as orc  # noqa: E402
from poi_recommendation_models_b200 import synthetic  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


N, D, hid, H, seed = 3000, 64, 64, 128, 21
rng = np.random.default_rng(seed)
coords, region, R = synthetic.make_catalog(N, seed=seed)
sd = orc.init_state("region_distance", N, D, hid, R, 1, seed=seed + 1, style="trained")
BM = 700
hist = np.stack([rng.choice(N, H, replace=False) for _ in range(BM)]).astype(np.int64)
tgt = rng.integers(0, N, BM).astype(np.int64)
tgt[::3] = hist[::3, H // 2]
aux = orc.latlon_abs_diff(coords, tgt, hist)
dscore = np.random.default_rng(5).normal(size=BM)


def grads(B, pp, order=None):
    m = util.make_model("region_distance", sd, 0.5)
    m.pairs_precision = pp
    idx = np.arange(B) if order is None else order
    s = m.attention_network(dev(hist[idx]), dev(tgt[idx]), dev(region[hist[idx]]), dev(region[tgt[idx]]), dev(aux[idx]))
    (s * dev(dscore[idx]).float()).sum().backward()
    torch.cuda.synchronize()
    return {n: p.grad.detach().double().cpu().numpy() for n, p in m.named_parameters() if p.grad is not None}


def report(tag, a, b):
    out = {}
    for k in ("attn_layer1.weight", "attn_layer1.bias", "attn_layer2.weight", "embed_history.weight", "dist_layer.bias"):
        e = np.abs(a[k] - b[k])
        out[k.split(".")[0] + ("b" if k.endswith("bias") else "")] = f"{e.max() / np.abs(b[k]).max():.1e}"
    print(tag, out, flush=True)


for B in (128, 256, 296, 297, 300, 400, 512, 592, 700):
    report(f"B={B:4d} tc vs fp32 bwd:", grads(B, ("fp32", "tc")), grads(B, ("fp32", "fp32")))
# the same 512 rows in another order: a precision effect does not care which CTA / tile slot a row lands in, a tile-reuse bug does
perm = np.random.default_rng(0).permutation(512)
g_a, g_b = grads(512, ("fp32", "tc")), grads(512, ("fp32", "tc"), perm)
report("B= 512 tc, rows permuted vs not:", g_a, g_b)
g_c = grads(512, ("fp32", "tc"))
report("B= 512 tc, run twice:", g_a, g_c)
# rows 296.. alone (they were second tiles above) vs their contribution
g_hi = grads(216, ("fp32", "tc"), np.arange(296, 512))
g_hi_ref = grads(216, ("fp32", "fp32"), np.arange(296, 512))
report("rows 296..511 alone, tc vs fp32:", g_hi, g_hi_ref)

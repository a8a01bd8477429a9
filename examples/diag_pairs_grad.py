"""Diagnostic: gradient error of the four (forward, backward) kernel pairings of the pair scorer against the float64 oracle,
per parameter tensor, on the C3-shaped test case of tests/test_gpu_pairs_tc.py."""
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


def case(N, D, hid, B, H, seed):
    rng = np.random.default_rng(seed)
    coords, region, R = synthetic.make_catalog(N, seed=seed)
    sd = orc.init_state("region_distance", N, D, hid, R, 1, seed=seed + 1, style="trained")
    hist = np.stack([rng.choice(N, H, replace=False) for _ in range(B)]).astype(np.int64)
    tgt = rng.integers(0, N, B).astype(np.int64)
    tgt[::3] = hist[::3, H // 2]
    aux = orc.latlon_abs_diff(coords, tgt, hist)
    return sd, hist, tgt, region, aux


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


for (B, H, seed, dseed) in ((512, 128, 21, 5), (256, 128, 384, 7), (512, 128, 22, 5)):
    sd, hist, tgt, region, aux = case(3000, 64, 64, B, H, seed)
    dscore = np.random.default_rng(dseed).normal(size=B)
    ref_s, ref = orc.grads(sd, "region_distance", 0.5, torch.from_numpy(hist), torch.from_numpy(tgt), torch.from_numpy(region[hist]),
                           torch.from_numpy(region[tgt]), torch.from_numpy(aux), torch.from_numpy(dscore))
    ref = {k: v.numpy() for k, v in ref.items()}
    print(f"case B={B} H={H} seed={seed}: max|score|={float(ref_s.abs().max()):.3g}")
    for pp in (("tc", "tc"), ("tc", "fp32"), ("fp32", "tc"), ("fp32", "fp32")):
        m = util.make_model("region_distance", sd, 0.5)
        m.pairs_precision = pp
        s = m.attention_network(dev(hist), dev(tgt), dev(region[hist]), dev(region[tgt]), dev(aux))
        (s * dev(dscore).float()).sum().backward()
        out = {}
        for name, p in m.named_parameters():
            if name not in ref:
                continue
            r = np.asarray(ref[name], dtype=np.float64)
            g = np.zeros_like(r) if p.grad is None else p.grad.detach().cpu().double().numpy()
            e = np.abs(g - r)
            i = np.unravel_index(np.argmax(e), e.shape)
            out[name.replace(".weight", "")] = f"{e.max() / max(np.abs(r).max(), 1e-300):.1e}@{i[0]}"
        print("  fwd/bwd", pp, out)

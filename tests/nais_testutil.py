"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

from oracle import nais_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-4  # north_star: per-pair scores within 1e-4 relative (condition-aware)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def golden_sd(z, prefix):
    return {k[len(prefix):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix)}


def make_model(variant, sd, beta, device="cuda"):
    from poi_recommendation_models_b200 import model as M
    N = sd["embed_history.weight"].shape[0]
    hid = sd["attn_layer1.weight"].shape[0]
    lanes = 2 if orc.VARIANTS[variant]["dist"] == "latlon" else 0
    D = sd["attn_layer1.weight"].shape[1] - lanes
    R = sd["embed_region.weight"].shape[0] if "embed_region.weight" in sd else 1
    cls = M.CLASSES[variant]
    if variant == "basic":
        m = cls(N, D, hid, beta)
    elif variant == "region":
        m = cls(N, D, hid, beta, R)
    else:
        m = cls(N, D, hid, beta, R, 1)
    m.load_state_dict(sd, strict=True)
    return m.to(device).eval()


def call(m, variant, hist, tgt, hreg, treg, aux, pre_sigmoid=True):
    fn = m.attention_network if pre_sigmoid else m.forward
    if variant == "basic":
        return fn(hist, tgt)
    if variant == "region":
        return fn(hist, tgt, hreg, treg)
    if variant == "distance":
        return fn(hist, tgt, aux) if pre_sigmoid else fn(hist, tgt, hreg, treg, aux)
    return fn(hist, tgt, hreg, treg, aux)


def cond_err(got, ref64, scale):
    """max |got-ref| / max(|ref|, sum_h|w_h s_h|)"""
    got, ref64, scale = (np.asarray(t, dtype=np.float64) for t in (got, ref64, scale))
    return float(np.max(np.abs(got - ref64) / np.maximum(np.maximum(np.abs(ref64), scale), 1e-30)))


def oracle_user_scores(sd, variant, beta, coords, region, history, cand, dtype=torch.float64):
    """Oracle scores (and conditioning scale) of one user against explicit candidates, history items included
    (their own cell masked)."""
    history = np.asarray(history, dtype=np.int64)
    cand = np.asarray(cand, dtype=np.int64)
    B = len(cand)
    hist = torch.from_numpy(history)[None, :].expand(B, -1)
    hreg = torch.from_numpy(region[history])[None, :].expand(B, -1)
    treg = torch.from_numpy(region[cand])
    aux = None
    if orc.VARIANTS[variant]["dist"] == "latlon":
        aux = torch.from_numpy(orc.latlon_abs_diff(coords, cand, history[None, :].repeat(B, 0)))
    elif orc.VARIANTS[variant]["dist"] == "km":
        aux = torch.from_numpy(orc.dist_km(coords[cand][:, None, 0], coords[cand][:, None, 1], coords[history][None, :, 0],
                                           coords[history][None, :, 1]))  # float64 km, powerLaw.dist
    s, scale = orc.attention_network_with_scale(sd, variant, beta, hist, torch.from_numpy(cand), hreg, treg, aux, dtype=dtype)
    return s.numpy(), scale.numpy()


def lists_equal_outside_ties(got_ids, got_scores, ref_scores_by_id, k, tol=TOL):
    """got list must be a valid top-k of ref scores: every listed id's ref score >= (k-th best ref score) - tie tol,
    and the ref order is respected outside tie groups."""
    ref_sorted = np.sort(np.asarray(list(ref_scores_by_id.values()), dtype=np.float64))[::-1]
    kth = ref_sorted[min(k, len(ref_sorted)) - 1]
    for pos, i in enumerate(got_ids[:k]):
        r = ref_scores_by_id[int(i)]
        band = tol * max(abs(r), abs(kth), 1e-30)
        assert r >= kth - band, f"id {i} at rank {pos} has ref score {r} < kth {kth}"
        # rank consistency: the ref score at this rank differs from r by at most the tie band
        assert abs(ref_sorted[pos] - r) <= tol * max(abs(r), abs(ref_sorted[pos]), 1e-30) + 0.0, (pos, r, ref_sorted[pos])

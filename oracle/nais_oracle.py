"""TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the reference's NAIS hot path.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import
this module, and only as the checker / the CPU arm being timed.  The product package
(`poi_recommendation_models_b200`) never imports it; its ops raise if the CUDA extension is missing.

Parity pin: the reference ships no tests, golden vectors or fixtures, so this restatement is pinned
against *outputs of the reference itself*: `tests/golden/make_golden.py` imports the unmodified
`/root/reference/{model,batches,validation,eval_metrics,powerLaw}.py` (through `oracle/ref_shim.py`), runs them on
seeded inputs and commits the inputs+outputs as `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks this file
against those fixtures (and against the live reference whenever `/root/reference` is present).

Arithmetic is torch-on-CPU (the same ATen ops the reference calls) in float32 or float64.
Every function cites the reference lines it follows.  Nothing here is copied from the reference: the math is
re-derived in (b, h) index form, see SURVEY.md §3.3.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

# ----------------------------------------------------------------------------------------------------------------------
# Variants of the scorer (reference classes in model.py)
# ----------------------------------------------------------------------------------------------------------------------
#   name              reference class (model.py)                              region  dist lanes        scale
VARIANTS = {
    "basic": dict(cls="NAIS_basic", has_region=False, dist="none", scale=0.0),  # :8-97
    "region": dict(cls="NAIS_regionEmbedding", has_region=True, dist="none", scale=0.0),  # :99-187
    "region_distance": dict(cls="NAIS_region_distance_Embedding", has_region=True, dist="latlon", scale=100.0),  # :189-304
    "distance": dict(cls="NAIS_distance_Embedding", has_region=False, dist="latlon", scale=1000.0),  # :306-408
    "disentangled": dict(cls="NAIS_region_distance_disentangled_Embedding", has_region=True, dist="km", scale=0.0),  # :410-541
}


def init_state(variant: str, item_num: int, embed_size: int, hidden_size: int, region_num: int = 1,
               dist_embed_size: int = 1, seed: int = 0, style: str = "reference") -> Dict[str, torch.Tensor]:
    """Parameter set with the reference's state_dict keys and shapes (model.py:190-229 and siblings).

    style="reference": embeddings N(0, 0.01), Linear weights U(-1/sqrt(fan_in), ..), biases 0 (model.py:219-229).
    style="trained":   embeddings N(0, 0.3), non-zero biases N(0, 0.1) — weights that make scores move
                       (SURVEY.md §8c caveat: at the reference init every output is 0.5±3e-4).
    """
    g = torch.Generator().manual_seed(seed)
    v = VARIANTS[variant]
    D, hid = embed_size, hidden_size
    std = 0.01 if style == "reference" else 0.3

    def normal(*shape, s=std):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * s

    def linear_w(out_f, in_f):
        bound = 1.0 / math.sqrt(in_f)
        return (torch.rand(out_f, in_f, generator=g, dtype=torch.float32) * 2 - 1) * bound

    sd: Dict[str, torch.Tensor] = {}
    if variant in ("region", "region_distance"):
        dp = D // 2
        sd["embed_history.weight"] = normal(item_num, dp)
        sd["embed_target.weight"] = normal(item_num, dp)
        sd["embed_region.weight"] = normal(region_num, dp)
    else:
        sd["embed_history.weight"] = normal(item_num, D)
        sd["embed_target.weight"] = normal(item_num, D)
        if variant == "disentangled":
            sd["embed_region.weight"] = normal(region_num, D)
    if variant in ("region_distance", "disentangled"):
        sd["embed_distance.weight"] = normal(dist_embed_size, D, s=0.01)
    lanes = 2 if v["dist"] == "latlon" else 0
    sd["attn_layer1.weight"] = linear_w(hid, D + lanes)
    sd["attn_layer1.bias"] = torch.zeros(hid) if style == "reference" else normal(hid, s=0.1)
    sd["attn_layer2.weight"] = linear_w(1, hid)
    if v["dist"] == "latlon":
        sd["dist_layer.weight"] = linear_w(2, 2)
        sd["dist_layer.bias"] = torch.zeros(2) if style == "reference" else normal(2, s=0.1)
    if variant == "disentangled":
        sd["region_attn_layer1.weight"] = linear_w(hid, D)
        sd["region_attn_layer1.bias"] = torch.zeros(hid) if style == "reference" else normal(hid, s=0.1)
        sd["region_attn_layer2.weight"] = linear_w(1, hid)
    return sd


_SCALE_SINK: list = []  # when a list is pushed here, _beta_attention appends sum_h |w_h s_h| (conditioning scale)


def _beta_attention(a: torch.Tensor, sim: torch.Tensor, mask: torch.Tensor, beta: float) -> torch.Tensor:
    """exp / mask / beta-smoothed normaliser / weighted similarity (model.py:279-293).

    No max-subtraction: the reference exponentiates raw logits, and (sum E)^beta is not shift-invariant.
    """
    e = torch.exp(a) * mask.to(a.dtype)
    denom = torch.pow(e.sum(-1, keepdim=True), beta)
    terms = (e / denom) * sim
    if _SCALE_SINK:
        _SCALE_SINK[-1].append(terms.detach().abs().sum(-1))
    return terms.sum(-1)


def attention_network_with_scale(*args, **kwargs):
    """(score, scale) with scale[b] = sum over branches of sum_h |w_bh s_bh|: the magnitude the score's rounding error
    is relative to when the terms cancel (SURVEY.md §7 'tolerance must be condition-aware')."""
    sink: list = []
    _SCALE_SINK.append(sink)
    try:
        s = attention_network(*args, **kwargs)
    finally:
        _SCALE_SINK.pop()
    return s, sum(sink)


def attention_network(sd: Dict[str, torch.Tensor], variant: str, beta: float, hist: torch.Tensor, tgt: torch.Tensor,
                      hreg: Optional[torch.Tensor] = None, treg: Optional[torch.Tensor] = None,
                      aux: Optional[torch.Tensor] = None, dtype: torch.dtype = torch.float32,
                      l1_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Pre-sigmoid score[B] of (history row, target) pairs.

    region_distance: model.py:246-297; basic: :57-89; region: :144-180; distance: :355-401; disentangled: :467-534.
    `aux` is ll[B,H,2] = (|dlat|,|dlon|) degrees for the lat/lon variants, dist_km[B,H] for the disentangled one.
    `l1_scale` [B,H,hid] multiplies the first layer's output before the ReLU: with keep_mask/(1-p) it restates the
    train-mode `relu(drop(attn_layer1(x)))` of NAIS_basic / NAIS_regionEmbedding (model.py:71,162) for a GIVEN mask.
    """
    v = VARIANTS[variant]
    P = {k: t.to(dtype) for k, t in sd.items()}
    q = P["embed_history.weight"][hist]  # [B,H,Dp]
    p = P["embed_target.weight"][tgt]  # [B,Dp]
    mask = hist != tgt[:, None]  # get_mask, model.py:299-302
    if variant == "disentangled":
        r_h = P["embed_region.weight"][hreg]
        r_t = P["embed_region.weight"][treg]
        bias = aux.to(dtype) * P["embed_distance.weight"][0].sum()  # model.py:497-501 (bucket 0 only)
        x = q * p[:, None, :]
        xr = r_h * r_t[:, None, :]
        a = torch.relu(x @ P["attn_layer1.weight"].T + P["attn_layer1.bias"]) @ P["attn_layer2.weight"][0] + bias
        ar = torch.relu(xr @ P["region_attn_layer1.weight"].T + P["region_attn_layer1.bias"]) @ P["region_attn_layer2.weight"][0] + bias
        return _beta_attention(a, x.sum(-1), mask, beta) + _beta_attention(ar, xr.sum(-1), mask, beta)
    if v["has_region"]:
        q = torch.cat((q, P["embed_region.weight"][hreg]), -1)
        p = torch.cat((p, P["embed_region.weight"][treg]), -1)
    x = q * p[:, None, :]  # [B,H,D]
    sim = x.sum(-1)
    if v["dist"] == "latlon":
        g = torch.sigmoid((aux.to(dtype) * v["scale"]) @ P["dist_layer.weight"].T + P["dist_layer.bias"])  # :265 / :366
        x = torch.cat((x, g), -1)
    t1 = x @ P["attn_layer1.weight"].T + P["attn_layer1.bias"]
    if l1_scale is not None:
        t1 = t1 * l1_scale.to(dtype)
    a = torch.relu(t1) @ P["attn_layer2.weight"][0]
    return _beta_attention(a, sim, mask, beta)


def forward(sd, variant, beta, hist, tgt, hreg=None, treg=None, aux=None, dtype=torch.float32) -> torch.Tensor:
    """sigmoid(attention_network(...)) — model.py:231-244."""
    return torch.sigmoid(attention_network(sd, variant, beta, hist, tgt, hreg, treg, aux, dtype))


def bce_loss(pred: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    """nn.BCELoss() (mean, log clamped at -100) — model.py:209, run.py:251."""
    return torch.nn.functional.binary_cross_entropy(pred, label.to(pred.dtype))


# ----------------------------------------------------------------------------------------------------------------------
# Geography (run.py:47-54, powerLaw.py:7-21)
# ----------------------------------------------------------------------------------------------------------------------
def latlon_abs_diff(coords: np.ndarray, tgt: np.ndarray, hist: np.ndarray) -> np.ndarray:
    """ll[b,h,:] = |coords[tgt_b] - coords[hist_bh]| in float64, cast to float32 — the value `latlon_mat[t,h]`
    (run.py:47-54) holds, looked up as in run.py:239-247 / validation.py:108-118, without the O(N^2) table."""
    c = np.asarray(coords, dtype=np.float64)
    return np.abs(c[np.asarray(tgt)][:, None, :] - c[np.asarray(hist)]).astype(np.float32)


def dist_km(lat1, lon1, lat2, lon2) -> np.ndarray:
    """Great-circle km by the spherical law of cosines with the 1e-6 short-circuit (powerLaw.py:7-21), vectorised."""
    lat1, lon1, lat2, lon2 = (np.asarray(t, dtype=np.float64) for t in (lat1, lon1, lat2, lon2))
    d2r = math.pi / 180.0
    phi1, phi2 = (90.0 - lat1) * d2r, (90.0 - lat2) * d2r
    c = np.sin(phi1) * np.sin(phi2) * np.cos((lon1 - lon2) * d2r) + np.cos(phi1) * np.cos(phi2)
    out = np.arccos(np.clip(c, -1.0, 1.0)) * 6371
    same = (np.abs(lat1 - lat2) < 1e-6) & (np.abs(lon1 - lon2) < 1e-6)
    return np.where(same, 0.0, out)


# ----------------------------------------------------------------------------------------------------------------------
# Batches (batches.py:67-139)
# ----------------------------------------------------------------------------------------------------------------------
def train_batch_region(indices: Sequence[int], num_poi: int, negative_num: int, region_of: np.ndarray, rng):
    """Per-user training batch (batches.py:67-108).  `rng` must offer .shuffle(list) (python `random` or a seeded
    `random.Random`).  Order: positives shuffled; negatives = shuffle(all - positives)[:|P|*num_ng];
    targets interleaved [p_i, n_i1..n_ik]; labels [1,0..0]; history = shuffled positives repeated per target."""
    positives = list(indices)
    rng.shuffle(positives)
    negative = list(set(range(num_poi)) - set(positives))
    rng.shuffle(negative)
    negative = np.array(negative[: len(positives) * negative_num]).reshape(-1, negative_num)
    tgt = np.concatenate((np.array(positives).reshape(-1, 1), negative), axis=-1).reshape(-1)
    B = len(tgt)
    label = np.tile(np.array([1.0] + [0.0] * negative_num, dtype=np.float32), len(positives))
    hist = np.tile(np.array(positives, dtype=np.int64), (B, 1))
    return hist, tgt.astype(np.int64), label, region_of[hist], region_of[tgt]


def test_candidates(history: Sequence[int], num_poi: int) -> np.ndarray:
    """All POIs the user has not visited in training, ascending (validation.py:86-87; CPython iterates a set of
    small ints built from range() in ascending order)."""
    m = np.ones(num_poi, dtype=bool)
    m[np.asarray(history, dtype=np.int64)] = False
    return np.nonzero(m)[0].astype(np.int64)


# ----------------------------------------------------------------------------------------------------------------------
# Full-rank evaluation (validation.py:62-131)
# ----------------------------------------------------------------------------------------------------------------------
@dataclass
class Catalog:
    coords: np.ndarray  # [N,2] float64 (lat, lon) — G.poi_coos / place_coords
    region: np.ndarray  # [N] int64 — businessRegionEmbedList (run.py:218-223)


def reference_model(variant: str, sd: Dict[str, torch.Tensor], beta: float):
    """An instance of the UNMODIFIED reference class of `variant` (model.py, imported through oracle/ref_shim.py from
    /root/reference or from the oracle/_ref snapshot) holding the weights `sd`, on the CPU in eval mode; None when no
    reference is reachable.  `fullrank_user(..., model=...)` then times / checks the reference's own forward."""
    from . import ref_shim
    if not ref_shim.reference_available():
        return None
    ref = ref_shim.load_reference("model")
    cls = getattr(ref, VARIANTS[variant]["cls"])
    N = sd["embed_history.weight"].shape[0]
    hid = sd["attn_layer1.weight"].shape[0]
    D = sd["attn_layer1.weight"].shape[1] - (2 if VARIANTS[variant]["dist"] == "latlon" else 0)
    R = sd["embed_region.weight"].shape[0] if "embed_region.weight" in sd else 1
    if variant == "basic":
        m = cls(N, D, hid, beta)
    elif variant == "region":
        m = cls(N, D, hid, beta, R)
    else:
        m = cls(N, D, hid, beta, R, sd["embed_distance.weight"].shape[0] if "embed_distance.weight" in sd else 1)
    m.load_state_dict(sd, strict=True)
    return m.cpu().eval()


def _reference_forward(model, variant, hist, tgt, hreg, treg, aux):
    if variant == "basic":
        return model(hist, tgt)
    if variant == "region":
        return model(hist, tgt, hreg, treg)
    return model(hist, tgt, hreg, treg, aux)


def fullrank_user(sd, variant, beta, cat: Catalog, history: Sequence[int], topk: int = 50, chunk: int = 2048,
                  dtype=torch.float32, return_all: bool = False, model=None):
    """One iteration of the user loop of NAIS_region_distance_validation (validation.py:84-127):
    candidates = all - history, scored in chunks of `chunk`, concatenated, torch.topk on post-sigmoid scores.
    `model`: a `reference_model(...)` — the chunks then go through the reference's own `forward` (the loop around it is
    this restatement either way: the reference's validator wants a dense [N,N,2] latlon_mat, 25 GB at N = 40 000)."""
    history = np.asarray(history, dtype=np.int64)
    cand = test_candidates(history, len(cat.region))
    v = VARIANTS[variant]
    preds = []
    hist_t = torch.from_numpy(history)
    for s in range(0, len(cand), chunk):
        tg = cand[s:s + chunk]
        B = len(tg)
        hist = hist_t[None, :].expand(B, -1)
        hreg = torch.from_numpy(cat.region[history])[None, :].expand(B, -1)
        treg = torch.from_numpy(cat.region[tg])
        aux = None
        if v["dist"] == "latlon":
            aux = torch.from_numpy(latlon_abs_diff(cat.coords, tg, history[None, :].repeat(B, 0)))
        elif v["dist"] == "km":
            c = cat.coords
            aux = torch.from_numpy(dist_km(c[tg][:, None, 0], c[tg][:, None, 1], c[history][None, :, 0],
                                           c[history][None, :, 1]).astype(np.float32))
        if model is not None:
            with torch.no_grad():
                preds.append(_reference_forward(model, variant, hist, torch.from_numpy(tg), hreg, treg, aux))
        else:
            preds.append(forward(sd, variant, beta, hist, torch.from_numpy(tg), hreg, treg, aux, dtype))
    pred = torch.cat(preds)
    k = min(topk, len(cand))
    val, idx = torch.topk(pred, k)
    rec = cand[idx.numpy()]
    if return_all:
        return rec, val.numpy(), cand, pred.numpy()
    return rec, val.numpy()


def fullrank(sd, variant, beta, cat: Catalog, indptr: np.ndarray, indices: np.ndarray, topk: int = 50,
             chunk: int = 2048, dtype=torch.float32, users: Optional[Sequence[int]] = None) -> List[List[int]]:
    """recommended_list of validation.py:62-127 for the CSR train matrix (indptr, indices)."""
    users = range(len(indptr) - 1) if users is None else users
    out = []
    for u in users:
        rec, _ = fullrank_user(sd, variant, beta, cat, indices[indptr[u]:indptr[u + 1]], topk, chunk, dtype)
        out.append([int(i) for i in rec])
    return out


# ----------------------------------------------------------------------------------------------------------------------
# Metrics (eval_metrics.py:36-69)
# ----------------------------------------------------------------------------------------------------------------------
def precision_at_k(actual, predicted, k: int) -> float:
    """mean over ALL users of |pos ∩ rec[:k]| / k (eval_metrics.py:36-44).  Accumulated sequentially like the
    reference's `+=` loop: Python >= 3.12's built-in sum() is compensated and differs in the last bit."""
    tot = 0.0
    for a, p in zip(actual, predicted):
        tot += len(set(a) & set(p[:k])) / float(k)
    return tot / len(predicted)


def recall_at_k(actual, predicted, k: int) -> float:
    """mean over users with positives of |pos ∩ rec[:k]| / |pos| (eval_metrics.py:46-56)."""
    tot, n = 0.0, 0
    for a, p in zip(actual, predicted):
        a = set(a)
        if a:
            tot += len(a & set(p[:k])) / float(len(a))
            n += 1
    return tot / n


def hitrate_at_k(actual, predicted, k: int) -> float:
    """fraction of users with positives that have >=1 hit in rec[:k] (eval_metrics.py:58-69)."""
    tot, n = 0.0, 0
    for a, p in zip(actual, predicted):
        a = set(a)
        if a:
            tot += 1.0 if a & set(p[:k]) else 0.0
            n += 1
    return tot / n


def precision_at_k_per_sample(actual, predicted, k: int) -> float:
    """(number of entries of `predicted` that are in `actual`) / k (eval_metrics.py:29-34): no cut at k, repeats count."""
    n = 0
    for x in predicted:
        if x in actual:
            n += 1
    return n / (k + 0.0)


def apk(actual, predicted, k: int = 10) -> float:
    """Average precision at k (eval_metrics.py:70-101): sum over ranks i < k of hits(i)/(i+1) at first occurrences of
    relevant items, divided by min(|actual|, k); 0.0 when `actual` is empty."""
    if len(predicted) > k:
        predicted = predicted[:k]
    score, hits = 0.0, 0.0
    for i, x in enumerate(predicted):
        if x in actual and x not in predicted[:i]:
            hits += 1.0
            score += hits / (i + 1.0)
    if not actual:
        return 0.0
    return score / min(len(actual), k)


def mapk(actual, predicted, k: int = 10) -> float:
    """numpy mean of apk over users (eval_metrics.py:105-125)."""
    return float(np.mean([apk(a, p, k) for a, p in zip(actual, predicted)]))


def ndcg_at_k(actual, predicted, k: int) -> float:
    """Binary NDCG@k.  NOT in the reference (no NDCG anywhere); project-defined once, here and in
    the product alike: DCG = sum_{i<k} [rec_i in pos]/log2(i+2); IDCG = sum_{i<min(k,|pos|)} 1/log2(i+2); mean over
    users with positives."""
    tot, n = 0.0, 0
    for a, p in zip(actual, predicted):
        a = set(a)
        if a:
            dcg = sum(1.0 / math.log2(i + 2) for i, x in enumerate(p[:k]) if x in a)
            idcg = sum(1.0 / math.log2(i + 2) for i in range(min(k, len(a))))
            tot += dcg / idcg
            n += 1
    return tot / n


def evaluate(actual, predicted, k_list) -> Tuple[List[float], List[float], List[float]]:
    """(precision, recall, hit) lists over k_list — what evaluate_mp returns (eval_metrics.py:3-27), without the
    multiprocessing pools."""
    return ([precision_at_k(actual, predicted, k) for k in k_list],
            [recall_at_k(actual, predicted, k) for k in k_list],
            [hitrate_at_k(actual, predicted, k) for k in k_list])


# ----------------------------------------------------------------------------------------------------------------------
# Gradients (autograd through the restated forward; run.py:248-254)
# ----------------------------------------------------------------------------------------------------------------------
def grads(sd, variant, beta, hist, tgt, hreg, treg, aux, dscore: torch.Tensor, dtype=torch.float64):
    """d(sum(dscore*score))/d(param) for every parameter, by torch autograd on `attention_network`."""
    P = {k: t.detach().to(dtype).clone().requires_grad_(True) for k, t in sd.items()}
    s = attention_network(P, variant, beta, hist, tgt, hreg, treg, aux, dtype)
    (s * dscore.to(dtype)).sum().backward()
    return s.detach(), {k: (t.grad if t.grad is not None else torch.zeros_like(t)) for k, t in P.items()}


def train_step_bce(sd, variant, beta, hist, tgt, hreg, treg, aux, label, lr=0.01, state_sum=None, dtype=torch.float32):
    """One reference train step (run.py:248-254): forward, mean BCE, backward, dense Adagrad(lr, eps=1e-10, wd=0).
    Returns (loss, new_sd, new_state_sum)."""
    P = {k: t.detach().to(dtype).clone().requires_grad_(True) for k, t in sd.items()}
    pred = torch.sigmoid(attention_network(P, variant, beta, hist, tgt, hreg, treg, aux, dtype))
    loss = bce_loss(pred, label)
    loss.backward()
    state_sum = state_sum or {k: torch.zeros_like(t) for k, t in P.items()}
    new_sd, new_sum = {}, {}
    for k, t in P.items():
        gk = t.grad if t.grad is not None else torch.zeros_like(t)
        new_sum[k] = state_sum[k].to(dtype) + gk * gk
        new_sd[k] = (t.detach() - lr * gk / (new_sum[k].sqrt() + 1e-10))
    return float(loss.detach()), new_sd, new_sum

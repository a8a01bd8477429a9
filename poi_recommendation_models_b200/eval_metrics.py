"""Drop-in for the reference's `eval_metrics.py` (precision/recall/hit-rate @k, eval_metrics.py:3-69; the per-sample
precision and average-precision helpers, eval_metrics.py:29-34, 70-125).

Numbers are bit-identical to the reference functions: the per-user hit counts are the same integers and they are
accumulated as Python floats in the same user order with the same divisions.  `evaluate_mp` keeps the reference's
name and return value but does not fork three `multiprocessing.Pool`s per call.  `ndcg_at_k` is project-defined
(the reference has no NDCG) — binary gains, log2 discount, ideal = min(k, |positives|) hits.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple


def _hits(actual, predicted, topk) -> Tuple[List[int], List[int]]:
    hits, sizes = [], []
    for a, p in zip(actual, predicted):
        a = set(int(x) for x in a)
        hits.append(len(a & set(int(x) for x in p[:topk])))
        sizes.append(len(a))
    return hits, sizes


def precision_at_k(actual, predicted, topk):
    hits, _ = _hits(actual, predicted, topk)
    s = 0.0
    for h in hits:
        s += h / float(topk)
    return s / len(predicted)


def recall_at_k(actual, predicted, topk):
    hits, sizes = _hits(actual, predicted, topk)
    s, n = 0.0, 0
    for h, m in zip(hits, sizes):
        if m != 0:
            s += h / float(m)
            n += 1
    return s / n


def hitrate_at_k(actual, predicted, topk):
    hits, sizes = _hits(actual, predicted, topk)
    s, n = 0.0, 0
    for h, m in zip(hits, sizes):
        if m != 0:
            if h > 0:
                s += 1
            n += 1
    return s / n


def precision_at_k_per_sample(actual, predicted, topk):
    """One user's precision (eval_metrics.py:29-34): every entry of `predicted` found in `actual` counts — the list is
    NOT cut to `topk` and repeated entries count again, exactly like the reference; `topk` is only the denominator."""
    return sum(1 for place in predicted if place in actual) / (topk + 0.0)


def apk(actual, predicted, k=10):
    """Average precision at k of one ranked list (eval_metrics.py:70-101): a position scores hits_so_far / position when
    its item is relevant and has not appeared earlier in the list; normalised by min(|actual|, k); 0.0 without positives."""
    ranked = predicted[:k] if len(predicted) > k else predicted
    seen, hits, score = [], 0.0, 0.0
    for pos, item in enumerate(ranked, start=1):
        if item in actual and item not in seen:
            hits += 1.0
            score += hits / float(pos)
        seen.append(item)
    if not actual:
        return 0.0
    return score / min(len(actual), k)


def mapk(actual, predicted, k=10):
    """Mean of `apk` over users (eval_metrics.py:105-125); numpy mean like the reference, so the same float64."""
    import numpy as np
    return np.mean([apk(a, p, k) for a, p in zip(actual, predicted)])


def ndcg_at_k(actual, predicted, topk):
    s, n = 0.0, 0
    for a, p in zip(actual, predicted):
        a = set(int(x) for x in a)
        if a:
            dcg = sum(1.0 / math.log2(i + 2) for i, x in enumerate(p[:topk]) if int(x) in a)
            idcg = sum(1.0 / math.log2(i + 2) for i in range(min(topk, len(a))))
            s += dcg / idcg
            n += 1
    return s / n


def evaluate_mp(positive_list, recommended_list, k_list):
    precision = [precision_at_k(positive_list, recommended_list, k) for k in k_list]
    recall = [recall_at_k(positive_list, recommended_list, k) for k in k_list]
    hit = [hitrate_at_k(positive_list, recommended_list, k) for k in k_list]
    return precision, recall, hit


def evaluate_device(positive_list, rec_ids, k_list):
    """(precision, recall, hit) lists like `evaluate_mp`, from a DEVICE tensor of recommended ids [U, k] (as
    `model.predict_topk` returns): the set overlaps are counted on the GPU (`nais_hits_at_k`), only U x len(k_list)
    integers come back, and the means are accumulated here exactly like eval_metrics.py:36-69 — same floats."""
    from . import ops
    hits = ops.hits_at_k(rec_ids, positive_list, k_list).cpu().tolist()
    sizes = [len(set(int(x) for x in p)) for p in positive_list]
    precision, recall, hit = [], [], []
    for i, k in enumerate(k_list):
        sp, sr, sh, n = 0.0, 0.0, 0.0, 0
        for u, m in enumerate(sizes):
            h = hits[u][i]
            sp += h / float(k)
            if m != 0:
                sr += h / float(m)
                if h > 0:
                    sh += 1
                n += 1
        precision.append(sp / len(sizes))
        recall.append(sr / n)
        hit.append(sh / n)
    return precision, recall, hit

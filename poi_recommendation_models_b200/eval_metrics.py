"""Drop-in for the reference's `eval_metrics.py` (precision/recall/hit-rate @k, eval_metrics.py:3-69).

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

"""Drop-in for the NAIS batch builders of the reference's `batches.py` (:67-139).

Same arguments, same outputs (int64 / float32 tensors on the CUDA device), same use of Python's global `random`
stream — so `random.seed(s)` gives the batch the reference would build — but the O(N) Python `set` difference is a
numpy mask.
"""
from __future__ import annotations

import random

import numpy as np
import torch


def _device():
    return torch.device("cuda" if torch.cuda.is_available() else "cpu")


def _non_visited(num_poi: int, visited) -> list:
    m = np.ones(num_poi, dtype=bool)
    m[np.asarray(visited, dtype=np.int64)] = False
    return np.nonzero(m)[0].tolist()  # ascending == CPython iteration order of set(range(N)) - set(visited)


def get_NAIS_batch_region(train_matrix, num_poi, uid, negative_num, businessRegionEmbedList):
    """Per-user training batch (batches.py:67-108): shuffled positives, `len(pos)*negative_num` negatives drawn as the
    head of a shuffle of all non-visited POIs, targets interleaved [p, n..n], labels [1, 0..0], history repeated."""
    dev = _device()
    region = np.asarray(businessRegionEmbedList)
    positives = train_matrix.getrow(uid).indices.tolist()
    random.shuffle(positives)
    negative = _non_visited(num_poi, positives)
    random.shuffle(negative)
    negatives = np.array(negative[:len(positives) * negative_num], dtype=np.int64).reshape(-1, negative_num)
    pos = np.array(positives, dtype=np.int64)
    data = np.concatenate((pos.reshape(-1, 1), negatives), axis=-1).reshape(-1)
    labels = np.tile(np.array([1.0] + [0.0] * negative_num, dtype=np.float32), len(positives))
    hist = np.broadcast_to(pos, (len(data), len(pos)))
    return (torch.from_numpy(np.ascontiguousarray(hist)).to(dev), torch.from_numpy(data).to(dev),
            torch.from_numpy(labels).to(dev), torch.from_numpy(np.ascontiguousarray(region[hist]).astype(np.int64)).to(dev),
            torch.from_numpy(region[data].astype(np.int64)).to(dev))


def get_NAIS_batch_test_region(train_matrix, uid, businessRegionEmbedList):
    """Per-user full-catalogue test batch (batches.py:110-139)."""
    dev = _device()
    region = np.asarray(businessRegionEmbedList)
    history = np.asarray(train_matrix.getrow(uid).indices, dtype=np.int64)
    negative = np.asarray(_non_visited(train_matrix.shape[1], history), dtype=np.int64)
    hist = np.broadcast_to(history, (len(negative), len(history)))
    label = torch.zeros(len(negative), dtype=torch.float32, device=dev)
    return (torch.from_numpy(np.ascontiguousarray(hist)).to(dev), torch.from_numpy(negative).to(dev), label,
            torch.from_numpy(np.ascontiguousarray(region[hist]).astype(np.int64)).to(dev),
            torch.from_numpy(region[negative].astype(np.int64)).to(dev))


def get_NAIS_batch(train_matrix, num_poi, uid, negative_num):
    """batches.py:24-50 (no regions)."""
    h, t, l, _, _ = get_NAIS_batch_region(train_matrix, num_poi, uid, negative_num, np.zeros(num_poi, dtype=np.int64))
    return h, t, l


def get_NAIS_batch_test(train_matrix, uid):
    """batches.py:52-65."""
    h, t, l, _, _ = get_NAIS_batch_test_region(train_matrix, uid, np.zeros(train_matrix.shape[1], dtype=np.int64))
    return h, t, l


def lat_lon_pairs(place_coords, target_pois, history_pois, device=None):
    """ll[b,h,:] = |coords[target_b] - coords[history_h]| as float32: the values the reference gathers from its dense
    float64 `latlon_mat` (run.py:47-54,239-247) without building the O(N^2) table."""
    c = np.asarray(place_coords, dtype=np.float64)
    t = np.asarray(target_pois, dtype=np.int64)
    h = np.asarray(history_pois, dtype=np.int64)
    if h.ndim == 1:
        h = np.broadcast_to(h, (len(t), len(h)))
    ll = np.abs(c[t][:, None, :] - c[h]).astype(np.float32)
    return torch.from_numpy(ll).to(device or _device())

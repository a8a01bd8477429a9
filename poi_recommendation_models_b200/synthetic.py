"""Seeded synthetic check-in data shaped like what the reference's `datasets.py` hands to `run.py`.

There is no network and the reference ships no data files, so every test and benchmark runs on data
from this generator.  It emits the same objects the reference's drivers consume (run.py:854):

* ``train`` — CSR matrix users x POIs (`Dataset.generate_data`, datasets.py:349-442), here as (indptr, indices);
* ``val_positive`` / ``test_positive`` — list of held-out POI ids per user;
* ``coords`` — float64 [N,2] (lat, lon), the `place_coords` / `G.poi_coos` table;
* ``region`` — int64 [N] dense region id per POI: a `size`-metre lat/lon grid over the bounding box
  (`get_region`, datasets.py:7-87) renumbered densely in ascending cell order (`get_region_num`, :146-181).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

_NYC = (40.55, 40.95, -74.05, -73.70)  # lat_min, lat_max, lon_min, lon_max
_NYC_N = 38333


def _haversine_m(lat1, lon1, lat2, lon2) -> float:
    r = 6371008.8
    p1, p2 = math.radians(lat1), math.radians(lat2)
    a = math.sin((p2 - p1) / 2) ** 2 + math.cos(p1) * math.cos(p2) * math.sin(math.radians(lon2 - lon1) / 2) ** 2
    return 2 * r * math.asin(math.sqrt(a))


def grid_regions(coords: np.ndarray, size_m: float = 300.0):
    """Vectorised `get_region` + `get_region_num` (datasets.py:7-87,146-181): rows = int(height/size),
    cols = int(mean(top,bottom width)/size), cell = row*cols+col (last row/col closed), then dense renumbering in
    ascending cell id.  Returns (region[N] int64, region_num)."""
    lat, lon = coords[:, 0], coords[:, 1]
    la0, la1, lo0, lo1 = lat.min(), lat.max(), lon.min(), lon.max()
    w1 = _haversine_m(la1, lo1, la1, lo0)
    w2 = _haversine_m(la0, lo1, la0, lo0)
    h1 = _haversine_m(la1, lo1, la0, lo1)
    cols = max(1, int((w1 + w2) / 2 / size_m))
    rows = max(1, int(h1 / size_m))
    r = np.minimum(((lat - la0) / ((la1 - la0) / rows)).astype(np.int64), rows - 1)
    c = np.minimum(((lon - lo0) / ((lo1 - lo0) / cols)).astype(np.int64), cols - 1)
    cell = r * cols + c
    uniq, dense = np.unique(cell, return_inverse=True)
    return dense.astype(np.int64), int(len(uniq))


@dataclass
class SyntheticCheckins:
    num_users: int
    num_pois: int
    coords: np.ndarray  # [N,2] float64
    region: np.ndarray  # [N] int64
    region_num: int
    indptr: np.ndarray  # [U+1] int64  (CSR train matrix)
    indices: np.ndarray  # [nnz] int64, ascending within a row like scipy CSR `.indices` after sort_indices
    val_positive: List[List[int]]
    test_positive: List[List[int]]

    def history(self, u: int) -> np.ndarray:
        return self.indices[self.indptr[u]:self.indptr[u + 1]]

    def train_csr(self):
        import scipy.sparse as sp
        data = np.ones(len(self.indices), dtype=np.float64)
        return sp.csr_matrix((data, self.indices, self.indptr), shape=(self.num_users, self.num_pois))


def make_catalog(num_pois: int, seed: int = 0, region_size_m: float = 300.0):
    """coords uniform in the NYC bounding box, scaled by sqrt(N/38333) per side so POI density stays constant."""
    rng = np.random.default_rng(seed)
    s = math.sqrt(num_pois / _NYC_N)
    la0, la1, lo0, lo1 = _NYC
    dlat, dlon = min((la1 - la0) * s, 60.0), min((lo1 - lo0) * s, 120.0)
    coords = np.empty((num_pois, 2), dtype=np.float64)
    coords[:, 0] = la0 + rng.random(num_pois) * dlat
    coords[:, 1] = lo0 + rng.random(num_pois) * dlon
    region, region_num = grid_regions(coords, region_size_m)
    return coords, region, region_num


def make_checkins(num_users: int, num_pois: int, hist_len: Optional[int] = None, seed: int = 0,
                  max_hist: int = 100, min_hist: int = 5, median_hist: float = 30.0, heldout: bool = True,
                  region_size_m: float = 300.0, local_frac: float = 0.7, local_sigma: float = 400.0) -> SyntheticCheckins:
    """`hist_len` fixed (configs C2-C5) or clipped log-normal sizes (C1: median≈30, max 100, min 5).

    Items: with prob `local_frac` a POI near the user's home in a spatially sorted order (discrete Gaussian offset,
    sigma `local_sigma` ranks), else a Zipf(1.0)-popular POI; without replacement.  Held-out: 20 % test / 10 % val
    of each user's items (at least one val), the split ratios of `train_test_val_split_with_time`
    (datasets.py:112-145).
    """
    rng = np.random.default_rng(seed + 1)
    coords, region, region_num = make_catalog(num_pois, seed, region_size_m)
    # spatial order: sort by (coarse lat band, lon) — neighbours in this order are near each other on the map
    band = ((coords[:, 0] - coords[:, 0].min()) / 0.01).astype(np.int64)
    order = np.lexsort((coords[:, 1], band))
    pop = 1.0 / np.arange(1, num_pois + 1)
    pop_cdf = np.cumsum(pop / pop.sum())
    pop_ids = rng.permutation(num_pois)

    if hist_len is None:
        sizes = np.clip(np.round(rng.lognormal(math.log(median_hist), 0.6, num_users)), min_hist, max_hist).astype(np.int64)
    else:
        sizes = np.full(num_users, hist_len, dtype=np.int64)
    totals = np.minimum(np.ceil(sizes / 0.7).astype(np.int64) + 1, num_pois) if heldout else sizes

    indptr = np.zeros(num_users + 1, dtype=np.int64)
    rows, vals, tests = [], [], []
    homes = rng.integers(0, num_pois, num_users)
    for u in range(num_users):
        need = int(totals[u])
        got: List[int] = []
        seen = set()
        while len(got) < need:
            m = 2 * (need - len(got)) + 8
            loc = np.clip(homes[u] + np.round(rng.normal(0, local_sigma, m)).astype(np.int64), 0, num_pois - 1)
            glob = pop_ids[np.minimum(np.searchsorted(pop_cdf, rng.random(m)), num_pois - 1)]
            pick = np.where(rng.random(m) < local_frac, order[loc], glob)
            for x in pick.tolist():
                if x not in seen:
                    seen.add(x)
                    got.append(x)
                    if len(got) == need:
                        break
        if heldout:
            n_train = int(sizes[u])
            rest = got[n_train:]
            n_val = max(1, len(rest) // 3)
            vals.append(rest[:n_val])
            tests.append(rest[n_val:])
            got = got[:n_train]
        else:
            vals.append([])
            tests.append([])
        rows.append(np.sort(np.asarray(got, dtype=np.int64)))
        indptr[u + 1] = indptr[u] + len(got)
    indices = np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64)
    return SyntheticCheckins(num_users, num_pois, coords, region, region_num, indptr, indices, vals, tests)

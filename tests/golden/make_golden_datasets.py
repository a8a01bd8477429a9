"""Generate tests/golden/datasets.npz by RUNNING THE UNMODIFIED REFERENCE `datasets.py` on a small synthetic dataset
directory (checkins.txt / poi_coos.txt written here in the reference's formats).

    python tests/golden/make_golden_datasets.py        # needs /root/reference

`datasets.py:6` imports the PyPI package `haversine` (not installed, no version pinned by the reference).  Its published
formula — great-circle distance on a sphere of mean radius 6371.0088 km — is supplied here as a stand-in module so that
`get_region` and `Dataset.read_poi_coos` run; it only decides how many grid rows / columns there are.
"""
from __future__ import annotations

import os
import random
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402


def _haversine_module():
    hv = types.ModuleType("haversine")
    R = 6371.0088

    def haversine(p1, p2, unit="km"):
        lat1, lng1, lat2, lng2 = map(np.radians, (p1[0], p1[1], p2[0], p2[1]))
        d = np.sin((lat2 - lat1) * 0.5) ** 2 + np.cos(lat1) * np.cos(lat2) * np.sin((lng2 - lng1) * 0.5) ** 2
        return float(2 * (R * (1000.0 if unit == "m" else 1.0)) * np.arcsin(np.sqrt(d)))

    def haversine_vector(a, b, unit="km", comb=False):
        a, b = np.radians(np.asarray(a, dtype=np.float64)), np.radians(np.asarray(b, dtype=np.float64))
        assert comb
        lat1, lng1 = a[:, 0][None, :], a[:, 1][None, :]
        lat2, lng2 = b[:, 0][:, None], b[:, 1][:, None]
        d = np.sin((lat2 - lat1) * 0.5) ** 2 + np.cos(lat1) * np.cos(lat2) * np.sin((lng2 - lng1) * 0.5) ** 2
        return 2 * R * np.arcsin(np.sqrt(d))

    hv.haversine, hv.haversine_vector = haversine, haversine_vector
    return hv


def write_dataset(path, U, N, seed):
    rng = np.random.default_rng(seed)
    lat = rng.uniform(40.70, 40.78, N)
    lng = rng.uniform(-74.02, -73.93, N)
    # POIs exactly on cell edges / the bounding box corners exercise the inclusive-edge rules of get_region
    lat[0], lng[0] = lat.min(), lng.min()
    lat[1], lng[1] = lat.max(), lng.max()
    lat[2], lng[2] = lat.max(), lng.min()
    order = rng.permutation(N)
    with open(os.path.join(path, "poi_coos.txt"), "w") as f:
        for lid in order:
            f.write(f"{lid} {float(lat[lid])!r} {float(lng[lid])!r}\n")
    lines = []
    for u in range(U):
        n = int(rng.integers(6, 40))
        pois = rng.choice(N, n, replace=False)
        for p in pois:
            for _ in range(int(rng.integers(1, 4))):
                lines.append((u, int(p), float(rng.integers(1_300_000_000, 1_400_000_000))))
    rng.shuffle(lines)
    with open(os.path.join(path, "checkins.txt"), "w") as f:
        for u, p, t in lines:
            f.write(f"{u}\t{p}\t{t}\n")
    return lat, lng


def main():
    sys.modules["haversine"] = _haversine_module()
    ref = ref_shim.load_reference("datasets")
    U, N = 24, 260
    out = {}
    with tempfile.TemporaryDirectory() as d:
        d = d + "/"
        lat, lng = write_dataset(d, U, N, seed=7)
        out["checkins_txt"] = np.frombuffer(open(d + "checkins.txt", "rb").read(), dtype=np.uint8)
        out["poi_coos_txt"] = np.frombuffer(open(d + "poi_coos.txt", "rb").read(), dtype=np.uint8)
        ds = ref.Dataset(U, N, d)
        raw, tm = ds.read_raw_data()
        random.seed(3)
        train, test_pos, val_pos = ds.split_data(raw, tm, 0)
        coords = ds.read_poi_coos(10)
        for name, m in (("raw", raw), ("time", tm), ("train", train)):
            m = m.tocsr()
            m.sort_indices()
            out[name + "_indptr"], out[name + "_indices"], out[name + "_data"] = m.indptr, m.indices, m.data
        out["test_flat"] = np.concatenate([np.asarray(t, dtype=np.int64) for t in test_pos])
        out["test_ptr"] = np.cumsum([0] + [len(t) for t in test_pos])
        out["val_flat"] = np.concatenate([np.asarray(t, dtype=np.int64) for t in val_pos])
        out["val_ptr"] = np.cumsum([0] + [len(t) for t in val_pos])
        out["place_coords"] = np.asarray(coords, dtype=np.float64)
        for size in (300, 1000):
            ref.get_region(coords, size, d)
            n = ref.get_region_num(d)
            out[f"cell_{size}"] = np.loadtxt(d + "poi_region.txt", dtype=np.int64)[:, 1]
            out[f"dense_{size}"] = np.loadtxt(d + "poi_region_sorted.txt", dtype=np.int64)[:, 1]
            out[f"region_num_{size}"] = np.int64(n)
            out[f"sorted_txt_{size}"] = np.frombuffer(open(d + "poi_region_sorted.txt", "rb").read(), dtype=np.uint8)
    out["U"], out["N"] = np.int64(U), np.int64(N)
    np.savez_compressed(os.path.join(HERE, "datasets.npz"), **out)
    print("wrote datasets.npz", {k: getattr(v, "shape", v) for k, v in out.items() if not k.endswith("_txt")})


if __name__ == "__main__":
    main()

"""Input-format loaders and the region grid (poi_recommendation_models_b200/datasets.py) against
tests/golden/datasets.npz — what the UNMODIFIED reference `datasets.py` produced from the same files
(tests/golden/make_golden_datasets.py).  Host-side numpy, no GPU."""
import os
import pickle
import sys

import numpy as np
import pytest
import scipy.sparse as sp

import nais_testutil as util
from poi_recommendation_models_b200 import datasets as D


@pytest.fixture(scope="module")
def golden(tmp_path_factory):
    z = util.load_golden("datasets.npz")
    d = tmp_path_factory.mktemp("ds")
    (d / "checkins.txt").write_bytes(z["checkins_txt"].tobytes())
    (d / "poi_coos.txt").write_bytes(z["poi_coos_txt"].tobytes())
    return z, str(d) + "/"


def _csr(z, name, shape):
    return sp.csr_matrix((z[name + "_data"], z[name + "_indices"], z[name + "_indptr"]), shape=shape)


def _same_csr(a, b):
    a, b = a.tocsr(), b.tocsr()
    a.sort_indices()
    b.sort_indices()
    return (np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices) and np.array_equal(a.data, b.data))


def test_read_checkins_and_poi_coos(golden):
    z, d = golden
    U, N = int(z["U"]), int(z["N"])
    raw, tm = D.read_checkins(d, U, N)
    assert _same_csr(raw, _csr(z, "raw", (U, N)))
    assert _same_csr(tm, _csr(z, "time", (U, N)))
    coords = D.read_poi_coos(d)
    assert np.array_equal(np.asarray(coords), z["place_coords"])  # file order (a dict in the reference), bit-exact floats


def test_split_with_time(golden):
    z, d = golden
    U, N = int(z["U"]), int(z["N"])
    train, test_pos, val_pos = D.split_with_time(_csr(z, "raw", (U, N)), _csr(z, "time", (U, N)))
    assert _same_csr(train, _csr(z, "train", (U, N)))
    for u in range(U):
        assert test_pos[u] == z["test_flat"][z["test_ptr"][u]:z["test_ptr"][u + 1]].tolist()
        assert val_pos[u] == z["val_flat"][z["val_ptr"][u]:z["val_ptr"][u + 1]].tolist()
    ds = D.Dataset(U, N, d)
    tr2, te2, va2, pc2 = ds.generate_data()
    assert _same_csr(tr2, train) and te2 == test_pos and va2 == val_pos and np.array_equal(np.asarray(pc2), z["place_coords"])


@pytest.mark.parametrize("size", [300, 1000])
def test_region_grid_and_dense_ids(golden, size, tmp_path):
    z, _ = golden
    cell, rownum, colnum = D.region_grid(z["place_coords"], size)
    assert np.array_equal(cell, z[f"cell_{size}"])  # incl. the POIs placed exactly on the bounding-box corners
    dense, n = D.densify_regions(cell)
    assert n == int(z[f"region_num_{size}"]) and np.array_equal(dense, z[f"dense_{size}"])
    assert D.write_region_files(str(tmp_path) + "/", cell) == n
    assert (tmp_path / "poi_region_sorted.txt").read_bytes() == z[f"sorted_txt_{size}"].tobytes()
    assert np.array_equal(D.read_region_list(str(tmp_path) + "/"), dense)


def test_region_grid_edge_points():
    """points on interior cell edges go to the upper cell, on the outer upper edges to the last row / column"""
    rng = np.random.default_rng(0)
    pts = np.stack([rng.uniform(10.0, 10.05, 500), rng.uniform(20.0, 20.06, 500)], 1)
    pts[0], pts[1] = (10.0, 20.0), (10.05, 20.06)
    cell, rownum, colnum = D.region_grid(pts, 500)
    alpha, delta = (10.05 - 10.0) / rownum, (20.06 - 20.0) / colnum
    pts2 = np.concatenate([pts, [[10.0 + alpha * 2, 20.0 + delta * 3], [10.05, 20.0 + delta * 1], [10.0 + alpha, 20.06]]])
    cell2, r2, c2 = D.region_grid(pts2, 500)
    assert (r2, c2) == (rownum, colnum) and np.array_equal(cell2[:500], cell)
    assert cell2[500] == 2 * colnum + 3 and cell2[501] == (rownum - 1) * colnum + 1 and cell2[502] == 1 * colnum + colnum - 1
    assert cell2[0] == 0 and cell2[1] == rownum * colnum - 1 and (cell2 >= 0).all()


def test_load_dataset_pickle_maps_reference_class(tmp_path, golden):
    """run.py:854 pickles an instance of the reference's `datasets.Dataset` inside the tuple"""
    z, d = golden
    import types
    mod = types.ModuleType("datasets")
    cls = type("Dataset", (), {})
    cls.__module__ = "datasets"
    mod.Dataset = cls
    sys.modules["datasets"] = mod
    try:
        obj = cls()
        obj.user_num, obj.poi_num, obj.directory_path = 5, 7, "./data/X/"
        payload = (sp.eye(3, format="csr"), [[1], [2]], [[0], [1]], [[1.0, 2.0]], obj)
        p = tmp_path / "dataset_X.pkl"
        with open(p, "wb") as f:
            pickle.dump(payload, f)
    finally:
        del sys.modules["datasets"]
    train, te, va, pc, ds = D.load_dataset_pickle(str(p))
    assert isinstance(ds, D.Dataset) and (ds.user_num, ds.poi_num, ds.directory_path) == (5, 7, "./data/X/")
    assert train.shape == (3, 3) and te == [[1], [2]] and pc == [[1.0, 2.0]]

"""TEST INFRASTRUCTURE ONLY — import shim for the *unmodified* reference modules.

The reference (`/root/reference/model.py:4-6`) imports three packages that are not installed in this image
(`torch_geometric`, `torchmetrics`, `haversine`).  None of them is used by any NAIS class on the hot path, so we
register empty stand-ins in `sys.modules` and import the reference modules from where they lie.  Nothing is copied.

Used only by `tests/golden/make_golden.py`, by CPU tests that are skipped when no reference is reachable, and by the CPU arm
of `bench.py` (`cpu_baseline.kind = "reference"`: the unmodified `model.py` classes, from /root/reference here or from the
`oracle/_ref/` snapshot on the GPU box).  The product package never imports this file.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

# where the unmodified modules are read from: the reference checkout if it exists (this container), else the snapshot that
# oracle/make_ref.py took of it (the GPU box has no /root/reference; oracle/_ref/ travels there, git-ignored)
_SNAPSHOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_DIR = os.environ.get("NAIS_REFERENCE_DIR") or ("/root/reference" if os.path.isfile("/root/reference/model.py") else _SNAPSHOT)


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "model.py"))


def _install_stubs() -> None:
    def _missing(*_a, **_k):  # pragma: no cover - never reached on the NAIS path
        raise RuntimeError("stubbed third-party symbol called; not part of the NAIS hot path")

    if "torch_geometric" not in sys.modules:
        tg = types.ModuleType("torch_geometric")
        tg_nn = types.ModuleType("torch_geometric.nn")
        tg_nn.GCNConv = type("GCNConv", (), {"__init__": _missing})
        tg.nn = tg_nn
        sys.modules["torch_geometric"] = tg
        sys.modules["torch_geometric.nn"] = tg_nn
    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")
        tm_f = types.ModuleType("torchmetrics.functional")
        tm_p = types.ModuleType("torchmetrics.functional.pairwise")
        tm_p.pairwise_manhattan_distance = _missing
        tm.functional = tm_f
        tm_f.pairwise = tm_p
        sys.modules["torchmetrics"] = tm
        sys.modules["torchmetrics.functional"] = tm_f
        sys.modules["torchmetrics.functional.pairwise"] = tm_p
    if "haversine" not in sys.modules:
        hv = types.ModuleType("haversine")
        hv.haversine = _missing
        hv.haversine_vector = _missing
        hv.Unit = types.SimpleNamespace(KILOMETERS="km", METERS="m")
        sys.modules["haversine"] = hv


def load_reference(*names: str):
    """Import reference modules by name (e.g. "model", "batches") under the prefix-free names they use themselves.

    They are inserted in ``sys.modules`` as ``_poi_ref_<name>`` as well so callers can tell them apart from the
    product package's own ``model`` mirror.
    """
    if not reference_available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_DIR}")
    _install_stubs()
    out = []
    added = False
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
        added = True
    try:
        for n in names:
            key = f"_poi_ref_{n}"
            if key in sys.modules:
                out.append(sys.modules[key])
                continue
            spec = importlib.util.spec_from_file_location(key, os.path.join(REFERENCE_DIR, n + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[key] = mod
            spec.loader.exec_module(mod)
            out.append(mod)
    finally:
        if added:
            sys.path.remove(REFERENCE_DIR)
    return out[0] if len(out) == 1 else tuple(out)

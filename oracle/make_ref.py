"""TEST INFRASTRUCTURE ONLY — recipe that snapshots the UNMODIFIED reference modules into `oracle/_ref/`.

The reference is pure Python: nothing to compile, so "building" it is a verbatim file snapshot of the five
modules on the hot path, taken from where they lie under /root/reference.  `oracle/_ref/` is git-ignored (never committed:
no reference source enters the history) but NOT gpurun-ignored, so the snapshot travels to the GPU box with the built .so
files; there `bench.py --impl reference` / `cpu_baseline` time the reference's own `model.py` classes on the host cores
(`cpu_baseline.kind = "reference"`).  Run by `__graft_entry__.build()` whenever /root/reference is present.

    python oracle/make_ref.py            # -> oracle/_ref/{model,batches,validation,eval_metrics,powerLaw}.py + MANIFEST.json
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("NAIS_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
MODULES = ("model", "batches", "validation", "eval_metrics", "powerLaw")


def snapshot() -> bool:
    if not os.path.isfile(os.path.join(SRC, "model.py")):
        return False
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for m in MODULES:
        src = os.path.join(SRC, m + ".py")
        shutil.copyfile(src, os.path.join(DST, m + ".py"))
        with open(src, "rb") as f:
            manifest[m + ".py"] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": manifest, "note": "verbatim, unmodified copies; never committed"}, f, indent=1)
    return True


if __name__ == "__main__":
    ok = snapshot()
    print("oracle/_ref written" if ok else f"{SRC} not present: nothing to snapshot")
    sys.exit(0)

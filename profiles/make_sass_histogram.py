#!/usr/bin/env python
"""profiles/r2_sass_histogram.txt: per-kernel SASS mnemonic counts of the shipped library (`cuobjdump -sass`), the evidence that the
hot kernels are tcgen05 / TMEM / bulk-copy code (UTCHMMA = tcgen05.mma kind::f16, UTCQMMA = kind::f8f6f4, LDTM / STTM = tcgen05.ld
/ st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk).  Run from the repo root after building: python profiles/make_sass_histogram.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "poi_recommendation_models_b200", "libnais_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "SYNCS", "FFMA", "MUFU", "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "RED", "ATOM"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    pat = re.compile(r"\b(" + "|".join(KEYS) + r")\b")
    rows, tot = [], collections.Counter()
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        c = collections.Counter(m.group(1) for m in pat.finditer(f))
        n = len(re.findall(r"^\s+/\*[0-9a-f]{4,6}\*/", f, flags=re.M))
        tot.update(c)
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        rows.append((dem[:108], n, c))
    out = [__doc__.strip().replace("\n", "\n# ").join(["# ", ""]), "",
           f"{'kernel':110s} {'instr':>7s} " + " ".join(f"{k:>7s}" for k in KEYS)]
    for name, n, c in sorted(rows, key=lambda r: -(r[2].get("UTCHMMA", 0) + r[2].get("UTCQMMA", 0)) * 100000 - r[1]):
        out.append(f"{name:110s} {n:7d} " + " ".join(f"{c.get(k, 0):7d}" for k in KEYS))
    out += ["", "TOTAL " + " ".join(f"{k}={tot.get(k, 0)}" for k in KEYS)]
    # the CTA-pair forms (tcgen05 ... cta_group::2) and LDGSTS (cp.async of the pair kernels' input pipeline), whole library
    mods = collections.Counter(re.findall(r"\b(UTC[A-Z]+(?:\.[A-Z0-9_]+)+|LDGSTS(?:\.[A-Z0-9_]+)*)", txt))
    out += ["MODIFIERS " + " ".join(f"{k}={v}" for k, v in sorted(mods.items(), key=lambda kv: -kv[1]))]
    with open(os.path.join(ROOT, "profiles", "r2_sass_histogram.txt"), "w") as fh:
        fh.write("\n".join(out) + "\n")
    print(out[-2])
    print(out[-1])


if __name__ == "__main__":
    main()

"""Where does a tile's time go in the two-threads-per-cell pair kernels?  Builds a PRIVATE copy of the library with
-DNAIS_PHASE_CLOCKS (thread 0 of every CTA accumulates clock64() deltas per phase; csrc/nais_pairs_tc.cu, nais_pairs_tc_bwd.cu),
runs the C3 training shape through it and prints the average clocks per tile and phase.  Diagnostic only: the product library
is never built with the flag.

  python examples/diag_phase_clocks.py --build        (here: cross-compile into examples/_phase_build/)
  gpurun -- python examples/diag_phase_clocks.py       (on the GPU box)
"""
import argparse
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "examples", "_phase_build")
LIB = os.path.join(OUT, "libnais_b200_phase.so")

FWD = ["prologue", "cp.async wait + barrier (a)", "rows x target, lat/lon gates", "barrier (b)", "P1 (next ids) + A image + barrier (c)",
       "MMA wait", "P2 (wait ids, request next rows)", "epilogue + pair barrier", "exp + barrier (d)", "row reduce + score"]
BWD = ["prologue", "cp.async wait + barrier (a)", "ids, gates, X image", "barrier (b)", "GEMM1 issue + P1 (next ids) + wait",
       "epilogue 1 + pair barrier", "dt image + dw2 shuffles + barrier (c)", "GEMM2/3 issue + 2nd gather loads + wait",
       "stage 2nd gather + epilogue 2 + wait ids", "barrier (d)", "P2 (request next rows) + dp reduce"]


def build():
    from poi_recommendation_models_b200 import build_ext as B
    os.makedirs(OUT, exist_ok=True)
    flags = [f for f in B.NVCC_FLAGS if f != "--use_fast_math=false"] + ["-DNAIS_PHASE_CLOCKS"]
    objs = []
    procs = []
    for src in B.SOURCES:
        obj = os.path.join(OUT, src.replace(".cu", ".o"))
        procs.append(subprocess.Popen([B._nvcc(), *flags, "-c", os.path.join(B.CSRC, src), "-o", obj]))
        objs.append(obj)
    for p in procs:
        if p.wait():
            raise SystemExit("nvcc failed")
    subprocess.check_call([B._nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
    for o in objs:
        os.remove(o)
    print(LIB)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--rows", type=int, default=8192)
    ap.add_argument("--hist", type=int, default=128)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    if args.build:
        return build()
    from poi_recommendation_models_b200 import _lib
    _lib.LIB_PATH = LIB  # before the first load
    import numpy as np
    import torch
    from poi_recommendation_models_b200 import model as M, synthetic
    _lib.load()
    dbg = C.CDLL(LIB)
    N, B_, H = 40000, args.rows, args.hist
    coords, region, R = synthetic.make_catalog(N, seed=0)
    torch.manual_seed(0)
    m = M.NAIS_region_distance_Embedding(N, 64, 64, 0.5, R, 1)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if name.startswith("embed_"):
                p.normal_(0, 0.3)
    m = m.cuda().train()
    g = torch.Generator(device="cuda").manual_seed(1)
    hist = torch.randint(0, N, (B_, H), device="cuda", generator=g)
    tgt = torch.randint(0, N, (B_,), device="cuda", generator=g)
    reg_t = torch.from_numpy(region).cuda()
    co = torch.from_numpy(coords).cuda()
    ll = (co[tgt][:, None, :] - co[hist]).abs().float().contiguous()
    hreg, treg = reg_t[hist], reg_t[tgt]
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    buf = (C.c_ulonglong * 16)()

    def run():
        m.zero_grad(set_to_none=True)
        s = m.attention_network(hist, tgt, hreg, treg, ll)
        (-torch.nn.functional.logsigmoid(s[: B_ // 2] - s[B_ // 2:]).mean()).backward()

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    dbg.nais_debug_phase_fwd(buf, 1)
    dbg.nais_debug_phase_bwd(buf, 1)
    for _ in range(args.iters):
        run()
    torch.cuda.synchronize()
    tiles = B_ * max(1, (H + 127) // 128) if H >= 128 else (B_ + (128 // H) - 1) // (128 // H)
    for name, fn, labels, per_sm in (("forward (3 CTAs/SM)", dbg.nais_debug_phase_fwd, FWD, 3), ("backward (2 CTAs/SM)", dbg.nais_debug_phase_bwd, BWD, 2)):
        fn(buf, 0)
        v = np.array(list(buf), dtype=np.float64)
        tot = v.sum()
        per_tile = v / (args.iters * tiles)
        print(f"\n{name}: {tot / (args.iters * tiles):.0f} clocks per tile per CTA; x tiles / ({sm} SMs x {per_sm}) = "
              f"{tot / args.iters / (sm * per_sm):.0f} clocks per launch")
        for i, lab in enumerate(labels):
            print(f"  {i:2d} {lab:46s} {per_tile[i]:9.0f}  {100 * v[i] / tot:5.1f} %")


if __name__ == "__main__":
    main()

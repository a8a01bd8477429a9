#!/usr/bin/env python
"""Headline benchmark: full-rank evaluation throughput of the NAIS region/distance scorer (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision tc_split|tc_fast|fp32]
    torchrun --nproc-per-node N ... bench.py --gpus N ...        (one rank per GPU, NCCL)

One *step* = one fused full-rank pass (scoring + top-k) of a batch of `--users-per-step` users against the whole
synthetic catalogue of the workload:

    C2 (default)  Gowalla-shaped: 30,000 users x 40,000 POIs, history 128, D = hid = 64, top-20   [BASELINE configs[1]]
    C4            Yelp-scale:     100,000 users x 1,000,000 POIs, history 128, D = hid = 64, top-20

Successive steps take successive user batches (wrapping around).  With N > 1 GPUs the catalogue is range-sharded
across ranks as far as a shard keeps >= 32k POIs (C4: 8 shards, as north_star / SURVEY.md §8e prescribe) and the user
batch is sliced across the remaining factor (C2's 40k POIs: user slices only; `distributed.grid_shape`); the per-rank
top-k lists are all-gathered over NCCL and, with > 1 shard, merged on device; total work per step is fixed ->
"scaling": "strong".

value   users/s with inputs resident in HBM (pair-scores/s = users/s x POIs is reported alongside)
e2e     the same metric through the drop-in API (`model.predict_topk` on host CSR arrays): pinned host -> device
        copy of the step's histories and device -> host copy of the top-k lists inside the timed region
roofline  tensor roofline of the dominant kernel: algorithmic FLOPs (F = 2*hid*(D+2) + 4*hid + 3*D + 16
        per (history item, candidate) cell) / CUDA-event time, against MEASURED_PEAKS.json bf16 (sustained)
cpu_baseline  the oracle port of validation.py:84-127 (torch CPU, all host threads) on a bounded sample of users

`--impl reference` times that CPU arm alone and prints the same JSON line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "C2": dict(users=30000, pois=40000, hist=128, D=64, hid=64, k=20, desc="Gowalla-shaped 30k users x 40k POIs, H=128, D=hid=64, top-20"),
    "C4": dict(users=100000, pois=1000000, hist=128, D=64, hid=64, k=20, desc="Yelp-scale 100k users x 1M POIs, H=128, D=hid=64, top-20"),
    "tiny": dict(users=512, pois=4096, hist=32, D=64, hid=64, k=20, desc="tiny self-test"),
}
BETA = 0.5


def flops_per_cell(D, hid):
    return 2 * hid * (D + 2) + 4 * hid + 3 * D + 16


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf_burst=d.get("bf16_tflops", 1590.0),
                    tf_sust=d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)), which="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, which="fallback")


def synth_histories(users, pois, hist, seed=0):
    """Fixed-length histories without replacement: popularity-skewed + local, cheap to generate for 100k users."""
    rng = np.random.default_rng(seed)
    homes = rng.integers(0, pois, users)
    out = np.empty((users, hist), dtype=np.int64)
    span = max(4 * hist, min(pois, 2000))
    for u0 in range(0, users, 4096):
        u1 = min(users, u0 + 4096)
        n = u1 - u0
        # local window around home (in id space, which the catalogue generator leaves spatially unordered -> mix)
        cand = (homes[u0:u1, None] + rng.integers(-span, span, (n, 3 * hist))) % pois
        glob = (rng.pareto(1.0, (n, 3 * hist)) * 50).astype(np.int64) % pois
        pick = np.where(rng.random((n, 3 * hist)) < 0.7, cand, glob)
        for i in range(n):
            _, first = np.unique(pick[i], return_index=True)
            sel = pick[i][np.sort(first)][:hist]
            if len(sel) < hist:  # top up
                extra = np.setdiff1d(rng.permutation(pois)[:4 * hist], sel)[:hist - len(sel)]
                sel = np.concatenate([sel, extra])
            out[u0 + i] = np.sort(sel)
    return out


class ClockSampler:
    """`nvidia-smi -lms 100` streamed for the duration of the timed region (the recipe's clocks line)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.3)  # first sample is in flight before the timed region starts
        except Exception:
            self.proc = None

    def stop(self):
        rows = []
        if self.proc is not None:
            try:
                time.sleep(0.15)
                self.proc.terminate()
                out, _ = self.proc.communicate(timeout=5)
                rows = [[x.strip() for x in ln.split(",")] for ln in out.splitlines() if ln.strip()]
            except Exception:
                pass
        sm, mx, pw, reasons = [], [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # "under load" = samples drawing more than half of the maximum power seen
        load = [c for c, p_ in zip(sm, pw) if pw and p_ >= 0.5 * max(pw)] or sm
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_under_load": len(load),
                "power_w_max": max(pw) if pw else None}


def cpu_arm(cfg, n_users, seed=0, warm=1, keep=None):
    """The reference's CPU path (oracle port of validation.py:84-127: chunks of 2048, torch.cat, torch.topk(50)) on the
    host cores, all threads.  Returns users/s over `n_users` users after `warm` warm-up users.  `keep` (a dict) receives the
    weights, inputs and the oracle's lists / full score vectors of the timed users, for `parity_block`."""
    import torch
    from oracle import nais_oracle as orc  # CPU arm only
    from poi_recommendation_models_b200 import synthetic
    torch.set_num_threads(os.cpu_count())
    coords, region, R = synthetic.make_catalog(cfg["pois"], seed=seed)
    sd = orc.init_state("region_distance", cfg["pois"], cfg["D"], cfg["hid"], R, 1, seed=seed + 1, style="trained")
    hist = synth_histories(warm + n_users, cfg["pois"], cfg["hist"], seed=seed + 2)
    cat = orc.Catalog(coords, region)
    with torch.no_grad():
        for u in range(warm):
            orc.fullrank_user(sd, "region_distance", BETA, cat, hist[u], 50)
        outs = []
        t0 = time.perf_counter()
        for u in range(warm, warm + n_users):
            outs.append(orc.fullrank_user(sd, "region_distance", BETA, cat, hist[u], 50, return_all=keep is not None))
        dt = time.perf_counter() - t0
    if keep is not None:
        keep.update(sd=sd, coords=coords, region=region, R=R, hist=hist[warm:warm + n_users], outs=outs)
    return n_users / dt, dt, torch.get_num_threads()


def parity_block(dev, cfg, kept, precision):
    """SURVEY.md §8(d) 'results-parity checks reported with every throughput number': the CPU arm's users scored again by the
    CUDA path with the SAME weights; our top-50 lists against the oracle's (validation.py:84-127 flow, fp32 torch CPU).
    Never raises: a failure is reported in the block, the throughput line stands."""
    try:
        import torch
        from poi_recommendation_models_b200 import model as M
        N, D, hid, H = cfg["pois"], cfg["D"], cfg["hid"], cfg["hist"]
        m = M.NAIS_region_distance_Embedding(N, D, hid, BETA, kept["R"], 1)
        m.load_state_dict(kept["sd"])
        m = m.to(dev).eval()
        m.set_catalog(region=kept["region"], coords=kept["coords"])
        hist = kept["hist"]
        n = len(hist)
        s, ids = m.predict_topk((np.arange(0, (n + 1) * H, H, dtype=np.int64), hist.reshape(-1)), 50, precision=precision)
        s, ids = s.cpu().numpy(), ids.cpu().numpy()
        exact, overlap, valid, rel = 0, 0, True, 0.0
        for u, (rec, val, cand, pred) in enumerate(kept["outs"]):
            exact += int([int(i) for i in rec] == ids[u].tolist())
            overlap += len(set(int(i) for i in rec) & set(ids[u].tolist()))
            by_id = dict(zip(cand.tolist(), pred.tolist()))
            mine = np.array([by_id[int(i)] for i in ids[u]], dtype=np.float64)  # the oracle's score of each POI we list
            kth = float(val[-1])
            valid = valid and bool((mine >= kth - 1e-4 * abs(kth)).all())
            rel = max(rel, float(np.max(np.abs(s[u].astype(np.float64) - val) / np.maximum(np.abs(val), 1e-30))))
        return {"users": n, "k": 50, "lists_identical": exact, "id_overlap": overlap / (50.0 * n),
                "valid_topk_of_oracle_scores_within_1e-4": valid, "max_rel_err_of_ranked_scores": rel,
                "against": "oracle port of validation.py:84-127 on the host (fp32 torch CPU), same weights and histories"}
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"}


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    n = max(1, args.cpu_users)
    t_all, vals = 0.0, []
    for _ in range(args.warmup):
        pass  # the CPU arm warms up inside cpu_arm (one untimed user per step)
    for s in range(args.steps):
        v, dt, threads = cpu_arm(cfg, n, seed=s)
        vals.append(v)
        t_all += dt
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": "fullrank_eval_users_per_sec", "value": value, "unit": "users/s",
            "pair_scores_per_sec": value * (cfg["pois"] - cfg["hist"]), "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * t_all / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config}: {cfg['desc']}", "users_per_step": n, "topk": 50},
            "cpu_baseline": {"value": value, "unit": "users/s", "cores": threads, "kind": "port",
                             "sample": f"{n} users x {cfg['pois']} POIs per step, {args.steps} steps, oracle port of validation.py:84-127 (chunk 2048, torch CPU)"},
            "e2e": {"value": value, "unit": "users/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_train(args, dev, lib, peaks, rank, world):
    """BASELINE config C3: BPR-style training forward + backward on (user, pos, neg) triples — 4,096 triples = 8,192
    (history row, target) pairs, history 128, D = hid = 64; every row has its own history, positives are in the history
    (live mask), negatives are not.  Step = pair forward, -log sigmoid(s+ - s-), hand-written backward of every
    parameter (attention MLP + embedding-row segment reduce).  Data parallel over triples with N > 1 (+ all-reduce)."""
    import torch
    import torch.distributed as dist
    from poi_recommendation_models_b200 import model as M, synthetic
    from poi_recommendation_models_b200.distributed import allreduce_gradients
    N, H, D, hid, T = 40000, 128, 64, 64, 4096
    coords, region, R = synthetic.make_catalog(N, seed=0)
    torch.manual_seed(1)
    m = M.NAIS_region_distance_Embedding(N, D, hid, BETA, R, 1)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if name.startswith("embed_"):
                p.normal_(0, 0.3)
    m = m.to(dev).train()
    hist_np = synth_histories(T, N, H, seed=3 + rank)
    rng = np.random.default_rng(4 + rank)
    pos = hist_np[np.arange(T), rng.integers(0, H, T)]
    neg = rng.integers(0, N, T)
    clash = (hist_np == neg[:, None]).any(1)
    neg[clash] = (neg[clash] + 1) % N
    hist = torch.from_numpy(np.concatenate([hist_np, hist_np])).to(dev)
    tgt = torch.from_numpy(np.concatenate([pos, neg])).to(dev)
    reg = torch.from_numpy(region).to(dev)
    c = torch.from_numpy(coords).to(dev)
    ll = (c[tgt][:, None, :] - c[hist]).abs().float().contiguous()
    hreg, treg = reg[hist], reg[tgt]

    def step():
        m.zero_grad(set_to_none=True)
        s = m.attention_network(hist, tgt, hreg, treg, ll)
        loss = -torch.nn.functional.logsigmoid(s[:T] - s[T:]).mean()
        loss.backward()
        allreduce_gradients(m, world)
        return loss

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = lib.nais_launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        loss = step()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    launches_timed = int(lib.nais_launch_count() - l0)  # this library's kernels inside the timed region only
    # the same step with the opt-in tensor-core pair kernels (NAIS_PAIRS_TC / NAIS_PAIRS_TC_BWD are read by the library per call)
    tc_opt = None
    if world == 1:
        try:
            os.environ["NAIS_PAIRS_TC"] = os.environ["NAIS_PAIRS_TC_BWD"] = "1"
            for _ in range(args.warmup):
                step()
            torch.cuda.synchronize()
            a3, b3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a3.record()
            for _ in range(args.steps):
                loss_tc = step()
            b3.record()
            torch.cuda.synchronize()
            tc_opt = {"ms_per_step": a3.elapsed_time(b3) / args.steps, "loss": float(loss_tc.detach()),
                      "note": "NAIS_PAIRS_TC=1 NAIS_PAIRS_TC_BWD=1: pair forward and backward contractions on tcgen05 (opt-in)"}
        except Exception as e:  # noqa: BLE001
            tc_opt = {"error": f"{type(e).__name__}: {e}"}
        finally:
            os.environ.pop("NAIS_PAIRS_TC", None)
            os.environ.pop("NAIS_PAIRS_TC_BWD", None)
    # the whole reference step (run.py:248-254: zero_grad, forward, BCELoss, backward, Adagrad.step) on 8192 pairs:
    # dense torch.optim.Adagrad over the [N, D/2] tables vs the row-sparse Adagrad fused into the segment reduce (f2),
    # at C3's catalogue (40k POIs: 5 MB tables) and at C4's (1M POIs: 128 MB tables, where the dense step is HBM traffic)
    full = None
    if world == 1:
        full = {"note": "8192 pairs, BCE, lr 0.01: whole optimizer step; the fused variant touches only the rows in the batch"}
        for n_big in (N, 1000000):
            if n_big == N:
                h2, t2, hr2, tr2, ll2, R2 = hist, tgt, hreg, treg, ll, R
            else:
                c2np, r2np, R2 = synthetic.make_catalog(n_big, seed=0)
                hn = synth_histories(T, n_big, H, seed=5)
                h2 = torch.from_numpy(np.concatenate([hn, hn])).to(dev)
                t2 = torch.from_numpy(np.concatenate([hn[:, 0], (hn[:, 1] + 1) % n_big])).to(dev)
                r2, c2 = torch.from_numpy(r2np).to(dev), torch.from_numpy(c2np).to(dev)
                ll2 = (c2[t2][:, None, :] - c2[h2]).abs().float().contiguous()
                hr2, tr2 = r2[h2], r2[t2]
            label = torch.cat([torch.ones(T), torch.zeros(T)]).to(dev)
            for kind in ("dense_adagrad", "fused_sparse_adagrad"):
                torch.manual_seed(2)
                m2 = M.NAIS_region_distance_Embedding(n_big, D, hid, BETA, R2, 1).to(dev).train()
                opt = torch.optim.Adagrad(m2.parameters(), lr=0.01, weight_decay=0.0)

                def full_step():
                    if kind == "dense_adagrad":
                        opt.zero_grad()
                        ls_ = m2.loss_func(m2(h2, t2, hr2, tr2, ll2), label)
                        ls_.backward()
                        opt.step()
                        return ls_.detach()
                    return m2.fused_adagrad_step(opt, label, h2, t2, hr2, tr2, ll2)

                for _ in range(args.warmup):
                    full_step()
                torch.cuda.synchronize()
                a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a2.record()
                for _ in range(args.steps):
                    ls = full_step()
                b2.record()
                torch.cuda.synchronize()
                full[f"{kind}_ms_N{n_big}"] = a2.elapsed_time(b2) / args.steps
                full[f"{kind}_loss_N{n_big}"] = float(ls)
                del m2, opt
    if rank == 0:
        cells = 2 * T * H
        F = flops_per_cell(D, hid)
        tf = cells * 4 * F * args.steps / (ms / 1000.0) / 1e12  # fwd F + bwd ~3F
        print(json.dumps({"metric": "bpr_train_triples_per_sec", "value": T * world * args.steps / (ms / 1000.0), "unit": "triples/s",
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                          "higher_is_better": True, "scaling": "weak", "dtype": "f32", "data": "synthetic",
                          "config": {"workload": "C3: 4096 (user,pos,neg) triples, H=128, D=hid=64, fwd+bwd (FP32 kernels)"},
                          "gpu_launches": launches_timed, "loss": float(loss.detach()), "tc_opt_in": tc_opt, "full_step": full,
                          "roofline": {"bound": "fp32-ffma", "achieved": tf, "unit": "TFLOP/s",
                                       "note": "algorithmic 4F per cell; CUDA-core path (tensor-core backward is next-round work)"}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=list(WORKLOADS))
    ap.add_argument("--precision", default=os.environ.get("NAIS_BENCH_PRECISION", "tc_auto"),
                    choices=["fp32", "tc_auto", "tc_split", "tc_mix", "tc_fast"])
    ap.add_argument("--users-per-step", type=int, default=0)
    ap.add_argument("--mode", default="eval", choices=["eval", "train"], help="eval = headline full-rank metric; train = C3 BPR fwd+bwd (triples/s)")
    ap.add_argument("--cpu-users", type=int, default=4, help="users in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = WORKLOADS[args.config]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return

    import torch
    import torch.distributed as dist
    from poi_recommendation_models_b200 import _lib, model as M, ops, synthetic
    from poi_recommendation_models_b200.distributed import ShardedRanker

    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    peaks = load_peaks()
    if args.mode == "train":
        run_train(args, dev, lib, peaks, rank, world)
        return

    U, N, H, D, hid, k = cfg["users"], cfg["pois"], cfg["hist"], cfg["D"], cfg["hid"], cfg["k"]
    ups = args.users_per_step or max(148, int({"fp32": 296, "tc_auto": 2368, "tc_split": 2368, "tc_mix": 2368, "tc_fast": 4736}[args.precision] * min(1.0, 40000 / N)))
    ups = min(ups, U)
    n_batches = min(args.steps + args.warmup, max(1, U // ups))
    # ---- synthetic data + random-init ("trained-like") weights of the named architecture ----------------------------
    coords, region, R = synthetic.make_catalog(N, seed=0)
    g = torch.Generator().manual_seed(1)
    torch.manual_seed(1)
    m = M.NAIS_region_distance_Embedding(N, D, hid, BETA, R, 1)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if name.startswith("embed_"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.3)
            elif name.endswith(".bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    m = m.to(dev).eval()
    m.set_catalog(region=region, coords=coords)
    hist_np = synth_histories(n_batches * ups, N, H, seed=2)
    ranker = ShardedRanker(m, rank, world)
    batches_dev, batches_host = [], []
    for b in range(n_batches):
        h = hist_np[b * ups:(b + 1) * ups]
        indptr = np.arange(0, (ups + 1) * H, H, dtype=np.int64)
        batches_dev.append(m.make_users(indptr, h.reshape(-1)))
        batches_host.append((torch.from_numpy(indptr).pin_memory(), torch.from_numpy(h.reshape(-1).copy()).pin_memory()))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_device(i):
        return ranker.topk(batches_dev[i % n_batches], k, precision=args.precision)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing ----------------------------------------------------------------------------------------
    for i in range(args.warmup):
        step_device(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = lib.nais_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (untimed)
        ev[i][0].record()
        out = step_device(args.warmup + i)
        ev[i][1].record()
    barrier()
    launches = lib.nais_launch_count() - l0
    clocks = sampler.stop()
    ms = [a.elapsed_time(b) for a, b in ev]
    t_local = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
    kern_ms = ranker.last_kernel_ms  # CUDA-event time of the dominant (scoring) kernel launch of the last step
    if world > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    total_ms = float(t_local.item())
    users_per_s = ups * args.steps / (total_ms / 1000.0)

    # ---- end to end through the drop-in API: host CSR in, host top-k out --------------------------------------------
    def step_e2e(i):
        ip, it = batches_host[i % n_batches]
        s, ids = ranker.topk_host(ip, it, k, precision=args.precision)
        return s, ids

    for i in range(min(args.warmup, 2)):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        s_h, i_h = step_e2e(args.warmup + i)
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_users_per_s = ups * args.steps / float(t_e2e.item())
    h2d = batches_host[0][0].numel() * 8 + batches_host[0][1].numel() * 8
    d2h = ups * k * (4 + 8)

    eff, choice = args.precision, None
    if args.precision == "tc_auto":  # the device-side gate picked MIX or SPLIT (no host sync inside the timed regions)
        choice = ops.last_tc_choice()
        eff = "tc_mix" if choice and choice["use_mix"] else "tc_split"
    if rank == 0:
        u0_, u1_, _ = ranker.user_range(ups)
        cells_per_launch = (u1_ - u0_) * H * (ranker.hi - ranker.lo)  # rank 0's launch: its user slice x its catalogue range
        F = flops_per_cell(D, hid)
        ach = cells_per_launch * F / (kern_ms / 1000.0) / 1e12 if kern_ms else None
        peak = peaks["tf_sust"]
        line = {"metric": "fullrank_eval_users_per_sec", "value": users_per_s, "unit": "users/s",
                "pair_scores_per_sec": users_per_s * N, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": {"fp32": "f32", "tc_split": "f16x2-split/f32-accum", "tc_mix": "f16+e5m2-corrections/f32-accum",
                          "tc_fast": "f16/f32-accum"}[eff],
                "data": "synthetic",
                "config": {"workload": f"{args.config}: {cfg['desc']}", "users_per_step": ups, "precision": args.precision,
                           "precision_effective": eff, "tc_auto_gate": choice,
                           "l2": "flushed between timed steps (256 MiB write)", "weights": "random init, trained-like scale",
                           "parallelism": ranker.parallelism if world > 1 else "single GPU"},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_users_per_s, "unit": "users/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                             "frac": (ach / peak) if ach else None, "traffic": None,
                             "kernel": ranker.kernel_name(args.precision), "kernel_ms": kern_ms,
                             "flops_per_cell": F, "cells_per_launch": cells_per_launch,
                             "peak_source": f"{peaks['which']} bf16 sustained (MEASURED_PEAKS.json)",
                             "hbm_frac": None}}
        if args.precision != "fp32":
            # FLOPs the tensor pipe actually executes per step of 256 cells (128 candidates x 2 history items):
            # SPLIT 3 passes x (D/16 + 1 ext) MMAs of 128 x nrow x 16; FAST 1 such pass + 2 passes of N = 16 (S/L rows)
            nrow = max((2 * hid + 4 + 15) // 16 * 16, 2 * hid + 16)
            ks = D // 16 + 1
            # MIX: 1 fp16 pass (incl. ONE ext MMA) + 2 e5m2 passes of D/32 K=32 MMAs, each occupying the pipe like an fp16
            # K=16 MMA (measured: tests/umma_probe_f8.cu) -> counted as fp16-equivalent pipe FLOPs
            per_step = {"tc_split": 3 * ks * 2 * 128 * nrow * 16,
                        "tc_mix": (ks + 2 * (D // 32)) * 2 * 128 * nrow * 16,
                        "tc_fast": ks * 2 * 128 * nrow * 16 + 2 * ks * 2 * 128 * 16 * 16}[eff]
            issued = cells_per_launch / 256 * per_step / (kern_ms / 1000.0) / 1e12
            line["roofline"].update({"issued_tflops": issued, "issued_frac": issued / peak,
                                     "issued_note": "tensor-pipe FLOPs executed (fp16-equivalent issue slots) incl. split/correction passes, ext K-step and S/L rows"})
        tr = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
        if os.path.isfile(tr) and eff in ("tc_split", "tc_mix"):
            with open(tr) as f:
                t = json.load(f)
            t = t.get(eff, t if "dram_bytes_per_user" in t else None)  # one entry per precision mode
            if t:
                line["roofline"]["traffic"] = t["dram_bytes_per_user"] * ups + t.get("dram_bytes_const", 0)
                line["roofline"]["traffic_source"] = t["source"]
        # algorithmic HBM bytes of the launch: catalogue rows + history items + output
        alg_bytes = (ranker.hi - ranker.lo) * (D // 2 * 4 + 4 + 8) + (u1_ - u0_) * H * (4 + D // 2 * 4 + 4 + 8) + (u1_ - u0_) * k * 8
        if kern_ms:
            line["roofline"]["hbm_frac"] = alg_bytes / (kern_ms / 1000.0) / 1e9 / peaks["hbm"]
        if not args.no_cpu_baseline and world == 1:
            kept = {}
            v, dt, threads = cpu_arm(cfg, max(1, args.cpu_users), keep=kept)
            line["parity"] = parity_block(dev, cfg, kept, args.precision)
            line["cpu_baseline"] = {"value": v, "unit": "users/s", "cores": threads, "kind": "port",
                                    "sample": f"{args.cpu_users} users x {N} POIs (H={H}), {dt:.1f} s, oracle port of validation.py:84-127 (chunk 2048, torch CPU, top-50)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

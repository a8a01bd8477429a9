#!/usr/bin/env python
"""Where does a C1 epoch at the reference's schedule (one user per optimizer step, nais_train_users) spend its time: the host
enqueueing ~14 launches per user, or the GPU running the chain of dependent launches?  Prints the time until the library call
returns (host enqueue) next to the time until the device has finished."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poi_recommendation_models_b200 import batches as PB, model as M, synthetic  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    U, N, D, hid, num_ng, lr = 1083, 38333, 64, 64, 4, 0.01
    data = synthetic.make_checkins(U, N, seed=0, hist_len=None, max_hist=100, min_hist=5, median_hist=30)
    torch.manual_seed(0)
    m = M.NAIS_region_distance_Embedding(N, D, hid, 0.5, data.region_num, 1).to(dev).train()
    opt = torch.optim.Adagrad(m.parameters(), lr=lr, weight_decay=0.0)
    bt = PB.DeviceBatcher(data.train_csr(), data.region, data.coords, device=dev, seed=0)
    order = np.random.default_rng(0).permutation(U)
    m.train_users(opt, bt, order[:16], num_ng, seed=0)
    torch.cuda.synchronize()
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        m.train_users(opt, bt, order, num_ng, seed=1 + rep)
        b.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"epoch {rep}: call returned after {1e3 * (t1 - t0):.1f} ms (host enqueue), device done after {1e3 * (t2 - t0):.1f} ms "
              f"(events: {a.elapsed_time(b):.1f} ms) -> {U / (t2 - t0):.0f} users/s, {1e6 * (t2 - t0) / U:.0f} us per user", flush=True)


    # device time alone: a long-running blocker goes first, so the host has enqueued the whole epoch before the device starts it
    x = torch.randn(8192, 8192, device=dev)
    for rep in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        for _ in range(60):
            x @ x  # ~1 TFLOP each
        a.record()
        m.train_users(opt, bt, order, num_ng, seed=11 + rep)
        b.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"behind a blocker: host enqueued everything after {1e3 * (t1 - t0):.1f} ms, device finished after {1e3 * (t2 - t0):.1f} ms; "
              f"device time of the epoch {a.elapsed_time(b):.1f} ms = {1e3 * a.elapsed_time(b) / U:.0f} us per user", flush=True)


if __name__ == "__main__":
    main()

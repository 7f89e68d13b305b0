#!/usr/bin/env python3
"""Where one slot's latency goes: device time per stage (the library's own CUDA events) and the wall clock of the host call,
for 1 / 2 / 8 slots of BASELINE config #1 (one message at -10 dB).  Exploratory; never a bench number."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ft8b200_loader import load

pkg = load()
dev = torch.device("cuda:0")
ctx = pkg.Context(0)
a = float(np.sqrt(2.0 * (2500.0 / 3200.0) * 10.0 ** (-10.0 / 10.0)))
for n in (1, 2, 8):
    sig = pkg.make_signals([(pkg.pack77_std("CQ", "K1JT", "FN20"), 700.0 + 50 * s, 0.5, a) for s in range(n)])
    d_i, d_q = ctx.synth_slots(sig, list(range(n + 1)), 1.0, 11)
    peak = torch.maximum(d_i.abs().amax(1), d_q.abs().amax(1))
    d_i = d_i * (0.5 / peak)[:, None]; d_q = d_q * (0.5 / peak)[:, None]
    h_i, h_q = d_i.cpu().numpy(), d_q.cpu().numpy()
    ctx.set_profiling(True)
    acc = {}
    for rep in range(12):
        ctx.process_slots(d_i, d_q); ctx.fetch_results(n)
        if rep >= 2:
            for k, v in ctx.stage_times().items():
                acc[k] = acc.get(k, 0.0) + max(v, 0.0) / 10
    ctx.set_profiling(False)
    t = []
    for rep in range(22):
        t0 = time.perf_counter()
        res, nres = ctx.process_slots_host(h_i, h_q)
        t.append((time.perf_counter() - t0) * 1e3)
    print("slots %d: host call %.3f ms (median), device stages us: %s, results %s" % (
        n, float(np.median(t[2:])), " ".join("%s=%.1f" % (k, v * 1e3) for k, v in acc.items() if v > 0), nres.tolist()), flush=True)

#!/usr/bin/env python3
"""Stand-alone time of each back-end stage through the stage-wise C ABI (CUDA events on the launching stream, whole GPU,
serial), for a small (128) and a large (4096) batch of 3200 sps slots, plus the worst case of the top-K selection
(noise waterfalls with min_score = 0: every position survives).  Writes gpurun_out/perf_kernels_<TAG>.json.
usage: tools/perf_kernels.py [TAG]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ft8b200_loader import load
from tools import synth

pkg = load()
dev = torch.device("cuda:0")
TAG = sys.argv[1] if len(sys.argv) > 1 else "x"
out = {}


def ev_time(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def make_slots(ctx, n, seed):
    rng = np.random.default_rng(seed)
    items = []
    for _ in range(n):
        to, de, ex = synth.random_message(rng)
        items.append((pkg.pack77_std(to, de, ex), float(rng.uniform(100.0, 1400.0)), float(rng.uniform(0.2, 0.8)), 0.28))
    sig = pkg.make_signals(items)
    d_i, d_q = ctx.synth_slots(sig, np.arange(n + 1, dtype=np.int32), 1.0, 7)
    return d_i, d_q, torch.maximum(d_i.abs().amax(1), d_q.abs().amax(1))


ctx = pkg.Context(0)
for n in (128, 4096):
    d_i, d_q, peak = make_slots(ctx, n, n)
    row = {}
    row["waterfall_ms"] = ev_time(lambda: ctx.waterfall(d_i, d_q, peak))
    mag = ctx.waterfall(d_i, d_q, peak)
    row["find_sync_ms"] = ev_time(lambda: ctx.find_sync(mag))
    cand, ncand = ctx.find_sync(mag)
    row["decode_ms"] = ev_time(lambda: ctx.decode(mag, cand, ncand))
    ok, stage, status, msg, _, _ = ctx.decode(mag, cand, ncand)
    row["spots_ms"] = ev_time(lambda: ctx.spots(cand, ncand, ok, msg, want_log=False))
    row["process_conditioned_ms"] = ev_time(lambda: ctx.process_conditioned(d_i, d_q, peak))
    row["candidates_per_slot"] = float(ncand.float().mean())
    row["waterfall_us_per_slot"] = row["waterfall_ms"] * 1e3 / n
    out["slots_%d" % n] = row
    print(n, row, flush=True)
    del d_i, d_q, mag
ctx.close()

# worst case of the selection: random waterfalls, min_score = 0 and -1000 (every one of the 35 856 positions reaches the heap stage)
rng = np.random.default_rng(5)
noise = torch.from_numpy(rng.integers(0, 256, size=(128, 94208), dtype=np.uint8)).to(dev)
for K, ms in ((120, 10), (120, 0), (120, -1000), (500, -1000)):
    c = pkg.Context(0, max_candidates=K, max_messages=50, min_score=ms)
    t = ev_time(lambda: c.find_sync(noise))
    out["find_sync_noise_K%d_min%d_ms" % (K, ms)] = t
    print("noise K", K, "min_score", ms, t, flush=True)
    c.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "perf_kernels_%s.json" % TAG), "w"), indent=1)

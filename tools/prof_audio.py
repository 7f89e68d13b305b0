#!/usr/bin/env python3
"""Short run of the 12 kHz monitor path (ft8b200_decode_audio: decode_ft8's main() batched) for ncu launch lists. Never a bench number."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ft8b200_loader import load
from tools import synth

pkg = load()
N = int(os.environ.get("PROF_SLOTS", "128"))
PER = int(os.environ.get("PROF_SIGNALS", "60"))
PROTO = int(os.environ.get("PROF_PROTOCOL", "1"))   # 1 = FT8 (15 s, 3840-point frames, 960 bins), 0 = FT4 (7.5 s, 1152-point frames, 288 bins)
rng = np.random.default_rng(4)
items, first = [], [0]
for s in range(N):
    for _ in range(PER):
        to, de, ex = synth.random_message(rng)
        items.append((pkg.pack77_std(to, de, ex), float(rng.uniform(200.0, 3000.0)), float(rng.uniform(0.0, 1.5)), float(rng.uniform(0.02, 0.5))))
    first.append(len(items))
ctx = pkg.Context(0)
aud = ctx.synth_audio(pkg.make_signals(items), first, PROTO, 0.05, 13, n_samples=180_000 if PROTO == 1 else 90_000)
torch.cuda.synchronize()
for rep in range(int(os.environ.get("PROF_REPS", "3"))):
    t0 = time.perf_counter()
    lines = pkg.decode_audio(ctx, aud, 12000, PROTO)
    torch.cuda.synchronize()
    print("rep", rep, "slots", N, "ms %.3f" % ((time.perf_counter() - t0) * 1e3), "decodes/slot %.1f" % np.mean([len(l) for l in lines]), flush=True)

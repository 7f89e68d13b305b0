#!/usr/bin/env python3
"""find_sync alone (score kernel + candidate selection) on synthetic 3200 sps slots, for ncu launch lists and captures.  Never a
bench number.
env: PROF_SLOTS (128), PROF_REPS (3), PROF_NOISE=1 adds random waterfalls with min_score 0."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ft8b200_loader import load
from tools import synth

pkg = load()
N = int(os.environ.get("PROF_SLOTS", "128"))
REPS = int(os.environ.get("PROF_REPS", "3"))
rng = np.random.default_rng(N)
items = []
for _ in range(N):
    to, de, ex = synth.random_message(rng)
    items.append((pkg.pack77_std(to, de, ex), float(rng.uniform(100.0, 1400.0)), float(rng.uniform(0.2, 0.8)), 0.28))
ctx = pkg.Context(0)
d_i, d_q = ctx.synth_slots(pkg.make_signals(items), np.arange(N + 1, dtype=np.int32), 1.0, 7)
peak = torch.maximum(d_i.abs().amax(1), d_q.abs().amax(1))
mag = ctx.waterfall(d_i, d_q, peak)
torch.cuda.synchronize()
for _ in range(REPS):
    cand, ncand = ctx.find_sync(mag)
torch.cuda.synchronize()
print("candidates/slot", float(ncand.float().mean()), flush=True)
if os.environ.get("PROF_NOISE"):
    noise = torch.from_numpy(np.random.default_rng(5).integers(0, 256, size=(N, 94208), dtype=np.uint8)).cuda()
    c = pkg.Context(0, max_candidates=120, max_messages=50, min_score=0)
    for _ in range(REPS):
        c.find_sync(noise)
    torch.cuda.synchronize()

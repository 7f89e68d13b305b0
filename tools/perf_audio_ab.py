#!/usr/bin/env python3
"""Config #3 on the 12 kHz monitor path (ft8b200_decode_audio) with both belief-propagation kernels, same inputs."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ft8b200_loader import load
from tools import synth
pkg = load()
def signals(rng, n, f_lo, f_hi, t_lo, t_hi, amp_lo, amp_hi):
    items = []
    for _ in range(n):
        to, de, ex = synth.random_message(rng)
        items.append((pkg.pack77_std(to, de, ex), float(rng.uniform(f_lo, f_hi)), float(rng.uniform(t_lo, t_hi)), float(rng.uniform(amp_lo, amp_hi))))
    return pkg.make_signals(items)
rng = np.random.default_rng(4)
NB = int(os.environ.get("PERF_AUDIO", "512"))
sigs = [signals(rng, 60, 200.0, 3000.0, 0.0, 1.5, 0.02, 0.5) for _ in range(NB)]
first = np.concatenate([[0], np.cumsum([s.size for s in sigs])]).astype(np.int32)
ctx = pkg.Context(0)
aud = ctx.synth_audio(np.concatenate(sigs), first, 1, 0.05, 13)
keep = None
for variant in (1, 0, 1, 0):
    pkg.set_decode_variant(variant)
    for _ in range(2):
        lines = pkg.decode_audio(ctx, aud, 12000, 1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        lines = pkg.decode_audio(ctx, aud, 12000, 1)
    torch.cuda.synchronize(); sec = (time.perf_counter() - t0) / 3
    blob = b"".join(l.tobytes() for l in lines)
    keep = keep or blob
    print("variant", variant, "ms", round(sec * 1e3, 3), "slots/s", round(NB / sec), "same", blob == keep, flush=True)
pkg.set_decode_variant(0)

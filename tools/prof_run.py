#!/usr/bin/env python3
"""Short whole-path run for ncu (launch list / --set full captures). Never a bench number."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ft8b200_loader import load

pkg = load()
B = int(os.environ.get("PROF_SLOTS", "16"))
reps = int(os.environ.get("PROF_REPS", "2"))
dev = torch.device("cuda:0")
if os.environ.get("PROF_INPUT", "bench") == "bench":
    import bench
    iq, _ = bench.gen_batch(B, 0, dev)   # one FT8 message per slot, like bench.py
else:
    g = torch.Generator(device=dev); g.manual_seed(1)
    iq = (torch.randn((B, pkg.RAW_SLOT_BYTES), device=dev, generator=g, dtype=torch.float16) * 30 + 127.5).clamp_(0, 255).to(torch.uint8)
ctx = pkg.Context(0)
for _ in range(reps):
    ctx.process_raw(iq, B)
torch.cuda.synchronize()
res, n = ctx.fetch_results(B)
print("slots", B, "n_results", n.tolist()[:8])

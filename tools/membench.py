#!/usr/bin/env python3
"""Read-only HBM bandwidth reference points (torch reductions) next to cic_block_sums_kernel. Exploratory."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ft8b200_loader import load
pkg = load()
dev = torch.device("cuda:0")
n = 32 * 72_000_000
x = torch.randint(0, 255, (n,), dtype=torch.uint8, device=dev)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
xi = x.view(torch.int32)
ms = t(lambda: xi.sum()); print(f"torch int32 sum  : {ms:.3f} ms  {n/ms/1e6:.0f} GB/s")
xf = x.view(torch.float32)
ms = t(lambda: torch.max(xi)); print(f"torch int32 max  : {ms:.3f} ms  {n/ms/1e6:.0f} GB/s")
y = torch.empty_like(x)
ms = t(lambda: y.copy_(x)); print(f"torch copy       : {ms:.3f} ms  {2*n/ms/1e6:.0f} GB/s (read+write)")
ctx = pkg.Context(0)
ctx.set_profiling(True)
xs = x.view(32, 72_000_000)
def k1():
    ctx.process_raw(xs, 32)
for _ in range(3): k1()
acc = 0
for _ in range(10):
    k1(); acc += ctx.stage_times()["block_sums"]
print(f"cic_block_sums   : {acc/10:.3f} ms  {n/(acc/10)/1e6:.0f} GB/s (input bytes only)")

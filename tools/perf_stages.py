#!/usr/bin/env python3
"""Per-stage device times (CUDA events recorded by the library) for a few batch shapes. Exploratory."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ft8b200_loader import load
import bench
from tools import synth, ft8enc

pkg = load()
dev = torch.device("cuda:0")
ctx = pkg.Context(0)
ctx.set_profiling(True)

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    acc = {}
    e0.record()
    for _ in range(reps):
        fn()
        for k, v in ctx.stage_times().items():
            acc[k] = acc.get(k, 0.0) + max(v, 0.0)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, {k: v / reps for k, v in acc.items()}

raw_sizes = [int(x) for x in os.environ.get("RAW_BATCHES", "32,96").split(",") if x]
if raw_sizes:
    B = max(raw_sizes)
    batch, _ = bench.gen_batch(B, 0, dev)
    for ov in (0, 2, 4):
        ctx.set_overlap(ov)
        for b in raw_sizes:
            ms, st = timeit(lambda: (ctx.process_raw(batch[:b], b), ctx.fetch_results(b)))
            print(f"raw  overlap={int(ov)} B={b:4d}: {ms:8.3f} ms/step  {b/ms*1e3:9.0f} slots/s  " + " ".join(f"{k}={v*1e3/b:6.2f}us" for k, v in st.items()))
    # two contexts on two streams: batch k+1's decimator overlaps batch k's back end
    ctx.set_overlap(0)
    ctx2 = pkg.Context(0)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for b in raw_sizes:
        def run(reps):
            for k in range(reps):
                c, s = (ctx, s1) if k % 2 == 0 else (ctx2, s2)
                with torch.cuda.stream(s):
                    c.process_raw(batch[:b], b)
                    c.fetch_results(b)
        run(4); torch.cuda.synchronize()
        t0 = time.time(); run(20); torch.cuda.synchronize(); dt = (time.time() - t0) / 20
        print(f"raw  2ctx(sync fetch) B={b:4d}: {dt*1e3:8.3f} ms/step {b/dt:9.0f} slots/s")
        def run2(reps):
            prev = None
            for k in range(reps):
                c, s = (ctx, s1) if k % 2 == 0 else (ctx2, s2)
                with torch.cuda.stream(s):
                    c.process_raw(batch[:b], b)
                if prev is not None:
                    with torch.cuda.stream(prev[1]):
                        prev[0].fetch_results(b)
                prev = (c, s)
            with torch.cuda.stream(prev[1]):
                prev[0].fetch_results(b)
        run2(4); torch.cuda.synchronize()
        t0 = time.time(); run2(20); torch.cuda.synchronize(); dt = (time.time() - t0) / 20
        print(f"raw  2ctx(pipelined fetch) B={b:4d}: {dt*1e3:8.3f} ms/step {b/dt:9.0f} slots/s")
    del batch
    torch.cuda.empty_cache()

# 3200 sps slots: crowded band (60 signals) and single signal, tiled
from oracle.pyoracle import Oracle
O = Oracle()
def tiled(kind, n):
    slots = []
    for s in range(8):
        if kind == "crowded":
            i_s, q_s, _ = synth.crowded_band(ft8enc, 60, 100 + s)
        else:
            i_s, q_s = synth.slot_f32([(ft8enc.tones(ft8enc.pack_std("CQ", "K1JT", "FN20")), 300.0 + 100 * s, 0.5, -10.0)], s)
        i_s, q_s, _ = O.condition(i_s, q_s, 48000)
        slots.append((i_s, q_s))
    hi = np.stack([s[0] for s in slots] * (n // 8)); hq = np.stack([s[1] for s in slots] * (n // 8))
    return torch.from_numpy(hi).to(dev), torch.from_numpy(hq).to(dev)
for kind in ("single", "crowded"):
    for n in [int(x) for x in os.environ.get("SLOT_BATCHES", "64,512,4096").split(",") if x]:
        d_i, d_q = tiled(kind, n)
        ms, st = timeit(lambda: (ctx.process_slots(d_i, d_q), ctx.fetch_results(n)), reps=3, warm=1)
        print(f"{kind:8s} N={n:5d}: {ms:8.3f} ms/step  {n/ms*1e3:9.0f} slots/s  " + " ".join(f"{k}={v*1e3/n:6.2f}us" for k, v in st.items() if v > 0))
        del d_i, d_q

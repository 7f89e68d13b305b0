#!/usr/bin/env python3
"""Sweep of the SM partition (ft8b200_pipe_set_partition): slots/s of the pipelined executor with the back end of batch n
on `back_sms` SMs and the decimator of batch n+1 on the rest, next to the serial and time-shared overlap modes. Exploratory."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ft8b200_loader import load
import bench

pkg = load()
dev = torch.device("cuda:0")
B = int(os.environ.get("RAW_BATCH", "256"))
depth = int(os.environ.get("DEPTH", "3"))
reps = int(os.environ.get("REPS", "12"))
parts = [int(x) for x in os.environ.get("PARTS", "0,-1,16,24,32,40,48,64").split(",") if x]
batch, _ = bench.gen_batch(B, 0, dev)
torch.cuda.synchronize()
ctx = pkg.Context(0)
ctx.process_raw(batch, B); ref_res, ref_n = ctx.fetch_results(B)
ctx.close()

def run(pipe, n):
    outs = []
    for k in range(n):
        if pipe.in_flight() == pipe.depth:
            outs.append(pipe.collect(B))
        pipe.submit(batch, B)
    while pipe.in_flight():
        outs.append(pipe.collect(B))
    return outs

for part in parts:
    pipe = pkg.Pipe(0, depth)
    try:
        if part == 0:
            pipe.set_mode(True); label = "serial"
        elif part < 0:
            pipe.set_mode(False); label = "overlap (time-shared)"
        else:
            f, b = pipe.set_partition(part); label = f"partition front={f} back={b}"
        outs = run(pipe, 4)
        same = all(np.array_equal(o[1], ref_n) and o[0].tobytes() == ref_res.tobytes() for o in outs)
        torch.cuda.synchronize()
        pipe.set_profiling(True)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0 = time.time()
        run(pipe, reps)
        torch.cuda.synchronize(); dt = (time.time() - t0) / reps
        st, nb = pipe.stage_times()
        tl = pipe.timeline()
        pipe.set_profiling(False)
        print(f"{label:32s} B={B} depth={depth}: {dt*1e3:8.3f} ms/step {B/dt:9.0f} slots/s identical={same}  " +
              " ".join(f"{k}={v/nb:6.3f}" for k, v in st.items()), flush=True)
        if os.environ.get("TIMELINE"):
            for row in tl[2:7]:
                print("    " + "  ".join(f"{n}[{a:7.3f},{b:7.3f}]" for n, (a, b) in zip(("k1", "fir", "wf", "sync", "dec", "spot"), row)), flush=True)
    except Exception as e:
        print(f"part={part}: FAILED {e}", flush=True)
    pipe.close()

#!/usr/bin/env python3
"""Pipelined executor (ft8b200_pipe_t) throughput for a few depths and batch sizes. Exploratory."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ft8b200_loader import load
import bench

pkg = load()
dev = torch.device("cuda:0")
sizes = [int(x) for x in os.environ.get("RAW_BATCHES", "32,64,128").split(",") if x]
depths = [int(x) for x in os.environ.get("DEPTHS", "1,2,3").split(",") if x]
B = max(sizes)
batch, _ = bench.gen_batch(B, 0, dev)
torch.cuda.synchronize()
ctx = pkg.Context(0)
ctx.process_raw(batch, B); ref_res, ref_n = ctx.fetch_results(B)
ctx.close()

def run(pipe, b, reps):
    outs = []
    for k in range(reps):
        if pipe.in_flight() == pipe.depth:
            outs.append(pipe.collect(b))
        pipe.submit(batch[:b], b)
    while pipe.in_flight():
        outs.append(pipe.collect(b))
    return outs

modes = [m for m in os.environ.get("MODES", "overlap,serial").split(",") if m]
for depth, mode in [(d, m) for d in depths for m in modes if not (d == 1 and m == "serial")]:
    pipe = pkg.Pipe(0, depth)
    pipe.set_mode(mode == "serial", int(os.environ.get("K1V", "-1")))
    for b in sizes:
        outs = run(pipe, b, 4)
        same = all(np.array_equal(o[1], ref_n[:b]) and o[0].tobytes() == ref_res[:b].tobytes() for o in outs)
        torch.cuda.synchronize()
        pipe.set_profiling(True)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record(); t0 = time.time()
        run(pipe, b, reps)
        e1.record(); torch.cuda.synchronize(); dt = (time.time() - t0) / reps
        st, nb = pipe.stage_times()
        pipe.set_profiling(False)
        print(f"pipe depth={depth} {mode:7s} B={b:4d}: {dt*1e3:8.3f} ms/step {b/dt:9.0f} slots/s identical={same}  " +
              " ".join(f"{k}={v/nb*1e3/b:6.2f}us" for k, v in st.items()))
    pipe.close()

# host input (e2e): pinned host batch
hb = int(os.environ.get("HOST_SLOTS", "8"))
host = torch.empty((hb, pkg.RAW_SLOT_BYTES), dtype=torch.uint8, pin_memory=True)
host.copy_(batch[:hb])
hnp = host.numpy()
for depth in depths:
    pipe = pkg.Pipe(0, depth)
    for per in (1, 2, 4, 8):
        if per > hb: continue
        def runh(reps):
            for k in range(reps):
                if pipe.in_flight() == pipe.depth:
                    pipe.collect(per)
                o = (k * per) % (hb - per + 1)
                pipe.submit_host(hnp[o:o + per], per)
            while pipe.in_flight():
                pipe.collect(per)
        runh(4)
        reps = 24
        t0 = time.time(); runh(reps); dt = (time.time() - t0) / reps
        print(f"host pipe depth={depth} slots/step={per}: {dt*1e3:8.3f} ms/step {per/dt:9.1f} slots/s  {per*72e6/dt/1e9:6.2f} GB/s H2D")
    pipe.close()

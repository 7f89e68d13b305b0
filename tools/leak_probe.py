#!/usr/bin/env python3
"""Which object leaves device memory behind when it is destroyed: free memory after N create/use/destroy cycles of each kind."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ft8b200_loader import load

pkg = load()
dev = torch.device("cuda:0")
rng = np.random.default_rng(1)
small = rng.integers(0, 256, size=12016 * 40, dtype=np.uint8)
d_raw = torch.from_numpy(small).to(dev)
aud = np.zeros(1920 * 4, np.float32)


def free():
    torch.cuda.synchronize(); torch.cuda.empty_cache()
    return torch.cuda.mem_get_info()[0]


def k_ctx():
    c = pkg.Context(0); c.close()
def k_ctx_used():
    c = pkg.Context(0); c.process_raw(d_raw, 1, small.size); c.fetch_results(1); c.close()
def k_stream():
    c = pkg.Context(0); st = pkg.Stream(c); st.callback(small[:65536]); st.flip(); st.fetch(); st.close(); c.close()
def k_pipe():
    p = pkg.Pipe(0, depth=3); p.close()
def k_pipe_used():
    p = pkg.Pipe(0, depth=3); p.submit(d_raw, 1, small.size); p.collect(1); p.close()
def k_pipe_host():
    p = pkg.Pipe(0, depth=3); p.submit_host(small, 1, small.size); p.collect(1); p.close()
def k_pipe_part():
    p = pkg.Pipe(0, depth=3); p.set_partition(32); p.close()
def k_pipe_part_used():
    p = pkg.Pipe(0, depth=3); p.set_partition(32); p.submit(d_raw, 1, small.size); p.collect(1); p.close()
def k_monitor():
    m = pkg.Monitor()
    for o in range(0, aud.size, 1920): m.process(aud[o:o + 1920])
    m.close()

for name, fn in list(globals().items()):
    if not name.startswith("k_"): continue
    fn(); fn()
    f0 = free()
    for _ in range(8): fn()
    f1 = free()
    print("%-18s %8.2f MiB per cycle" % (name[2:], (f0 - f1) / 8 / (1 << 20)), flush=True)

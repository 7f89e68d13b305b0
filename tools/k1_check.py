#!/usr/bin/env python3
"""cic_block_sums variants: bit-equality against the streaming kernel and stand-alone bandwidth. Exploratory."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ft8b200_loader import load
pkg = load()
dev = torch.device("cuda:0")
ctx = pkg.Context(0)
g = torch.Generator(device=dev); g.manual_seed(3)
variants = [int(x) for x in os.environ.get("VARIANTS", "1,2,3,4,5,6").split(",")]
for nstreams, nbytes in ((1, 12016 * 3), (3, 12016 * 40 + 751 * 2 * 4), (2, 72_000_000), (int(os.environ.get("BIG", "16")), 72_000_000)):
    stride = (nbytes + 15) // 16 * 16
    iq = torch.randint(0, 256, (nstreams, stride), dtype=torch.uint8, device=dev, generator=g)
    ctx.set_decimator_variant(0)
    ri, rq, rc, rp, _ = ctx.decimate(iq, nstreams, nbytes, stride)
    torch.cuda.synchronize()
    ri = ri.clone(); rq = rq.clone()
    for v in variants:
        ctx.set_decimator_variant(v)
        di, dq, dc, dp, _ = ctx.decimate(iq, nstreams, nbytes, stride)
        torch.cuda.synchronize()
        ok = torch.equal(di.view(torch.int32), ri.view(torch.int32)) and torch.equal(dq.view(torch.int32), rq.view(torch.int32))
        line = f"streams={nstreams} bytes={nbytes} variant={v} identical={ok}"
        if nbytes >= 72_000_000 and nstreams >= 8:
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            for vv in (0, v):
                ctx.set_decimator_variant(vv)
                for _ in range(3): ctx.decimate(iq, nstreams, nbytes, stride)
                e0.record()
                for _ in range(10): ctx.decimate(iq, nstreams, nbytes, stride)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                line += f"  v{vv}: {ms:.3f} ms ({nstreams*72.383488e6/ms/1e6:.0f} GB/s incl comb_fir)"
        print(line, flush=True)

#!/usr/bin/env python
"""Which SMs should the back-end partition take?  (run under gpurun; writes gpurun_out/part_layout_<tag>.json)

ft8b200_pipe_set_partition(back_sms + 1000 * layout): layout 0 is the driver's split by count, 1..6 compose the back partition of
8-SM groups spread over the driver's enumeration (csrc/pipe.cu).  For every (size, layout) this prints the hardware SM ids of the
back partition and the steady-state ms per 128-slot batch of the bench workload through ft8b200_pipe_autotune's probe, plus the
block-sum kernel's own launch time inside the partitioned executor (CUDA events around every launch)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from ft8b200_loader import load


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    sizes = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["32", "24", "40"])]
    layouts = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else "0,1,2,3,4,5,6".split(","))]
    pkg = load()
    device = torch.device("cuda", 0)
    Bc = 128
    batch, texts = bench.gen_batch(Bc, 0, device)
    torch.cuda.synchronize()
    pipe = pkg.Pipe(0, 3)
    pipe.set_mode(serial=False, decimator_variant=0)
    out = {"slots_per_batch": Bc, "points": []}
    for size in sizes:
        for lay in layouts:
            enc = size + 1000 * lay
            rec = {"back_sms_requested": size, "layout": lay}
            try:
                f, b = pipe.set_partition(enc)
                rec["front_sms"], rec["back_sms"] = f, b
                rec["back_smids"] = pipe.partition_smids(1)
                rec["front_smid_count"] = len(pipe.partition_smids(0))
                t = pipe.autotune(batch, Bc, candidates=(enc,), batches=48)
                rec["ms_per_batch"] = t["points"]
                # the block-sum kernel's own launch time inside the partitioned executor, comb+FIR on the back set
                pipe.set_profiling(True)
                n = 0
                for i in range(40):
                    while pipe.in_flight() >= 3:
                        pipe.collect(Bc); n += 1
                    pipe.submit(batch, Bc)
                while pipe.in_flight():
                    pipe.collect(Bc); n += 1
                ms, nb = pipe.stage_times()
                pipe.set_profiling(False)
                rec["stage_ms_per_launch"] = {k: v / max(nb, 1) for k, v in ms.items()}
                rec["comb_front_during_stage_times"] = t["comb_front"]
            except Exception as exc:
                rec["error"] = str(exc)
            print(json.dumps(rec), file=sys.stderr, flush=True)
            out["points"].append(rec)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/part_layout_%s.json" % tag, "w") as fh:
        json.dump(out, fh, indent=1)
    good = [r for r in out["points"] if "ms_per_batch" in r and r["ms_per_batch"]]
    good.sort(key=lambda r: min(r["ms_per_batch"].values()))
    for r in good[:8]:
        print(r["back_sms_requested"], r["layout"], r["ms_per_batch"], "k1 %.4f" % r["stage_ms_per_launch"]["block_sums"])


if __name__ == "__main__":
    main()

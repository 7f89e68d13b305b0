#!/usr/bin/env python3
"""Host->device copy ceiling of the box, with NO kernels: every rank (one process per GPU, bound to its GPU's NUMA node like
bench.py) loops cudaMemcpyAsync of a pinned 576 MB buffer (8 raw slots, the e2e measurement's step) into device memory, two
copies in flight; barrier + sync | CUDA events | barrier + sync, max over ranks.  What the e2e numbers of bench.py are bounded by.

  python tools/h2d_probe.py                                                   one GPU
  python -m torch.distributed.run --nproc-per-node N ... tools/h2d_probe.py   N GPUs
  tools/h2d_probe.py --merge out.json a.json b.json ...                       combine per-N lines into profiles/h2d_ceiling_r2.json's form
Prints one JSON line on rank 0."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 2 and sys.argv[1] == "--merge":
    rows = [json.loads(open(f).read().strip().splitlines()[-1]) for f in sys.argv[3:]]
    out = {"what": "pinned cudaMemcpyAsync loops, 576 MB per copy, 2 in flight per GPU, no kernels; one process per GPU", "unit": "GB/s aggregate",
           "by_gpus": {str(r["n_gpus"]): r["h2d_gbs"] for r in rows}, "per_gpu": {str(r["n_gpus"]): r["h2d_gbs"] / r["n_gpus"] for r in rows}, "rows": rows}
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print(json.dumps(out["by_gpus"]))
    sys.exit(0)

import torch
import torch.distributed as dist
import bench

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
numa = bench.bind_to_gpu_numa(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
BYTES = 8 * 72_000_000
host = [torch.empty(BYTES, dtype=torch.uint8).pin_memory() for _ in range(2)]
for h in host:
    h.fill_(7)
devb = [torch.empty(BYTES, dtype=torch.uint8, device=dev) for _ in range(2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]


def run(steps):
    for k in range(steps):
        with torch.cuda.stream(streams[k % 2]):
            devb[k % 2].copy_(host[k % 2], non_blocking=True)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


run(4)
steps = 24
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
barrier()
e0.record()
run(steps)
for s in streams:
    torch.cuda.current_stream().wait_stream(s)
e1.record()
barrier()
ms = e0.elapsed_time(e1)
t = torch.tensor([ms], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"n_gpus": world, "h2d_gbs": world * steps * BYTES / (float(t.item()) * 1e-3) / 1e9, "ms": float(t.item()), "steps": steps,
                      "bytes_per_copy": BYTES, "numa": numa, "host_cpus": os.cpu_count()}))
if world > 1:
    dist.destroy_process_group()

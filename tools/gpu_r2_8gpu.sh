#!/bin/bash
# Round-2 pass on one 8-GPU box (run under `gpurun --gpus 8`): host-copy ceiling at 1/2/4/8 GPUs, the multi-GPU library tests, the
# bench under torchrun at 8, and the single-process cluster arm at 8.  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python tools/h2d_probe.py > gpurun_out/h2d_1.json 2> gpurun_out/h2d_1.err
for N in 2 4 8; do
  $TR --nproc-per-node $N --master-port $((29600+N)) tools/h2d_probe.py > gpurun_out/h2d_$N.json 2> gpurun_out/h2d_$N.err
done
python tools/h2d_probe.py --merge gpurun_out/h2d_ceiling_r2.json gpurun_out/h2d_1.json gpurun_out/h2d_2.json gpurun_out/h2d_4.json gpurun_out/h2d_8.json
timeout 600 python -m pytest tests/test_gpu_cluster.py -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_cluster_8gpu.log; cat gpurun_out/pytest_cluster_8gpu.log
timeout 600 $TR --nproc-per-node 8 --master-port 29700 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r2_8gpu.json 2> gpurun_out/bench_r2_8gpu.err; echo bench8 rc=$?
timeout 600 python bench.py --gpus 8 --cluster --steps 10 --warmup 3 > gpurun_out/bench_r2_cluster8.json 2> gpurun_out/bench_r2_cluster8.err; echo cluster8 rc=$?
python - <<'PY'
import json
for f in ("bench_r2_8gpu", "bench_r2_cluster8"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, d["value"], d.get("verify"), (d.get("e2e") or {}).get("value"), (d.get("e2e_slots") or {}).get("value"))
        if "configs" in d:
            print({k: v.get("slots_per_s") for k, v in d["configs"].items()})
    except Exception as e:
        print(f, "FAILED", e)
PY

#!/bin/bash
# scaling pass on one multi-GPU box: bench.py at N = 2, 4, 8 (whatever the box has), launched the way the driver launches it
TAG=${1:-r1e}
mkdir -p gpurun_out
NG=$(python -c "import torch; print(torch.cuda.device_count())")
echo "gpus on the box: $NG"
for N in 2 4 8; do
  [ $N -le $NG ] || continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --cpu-slots 4 \
      > gpurun_out/bench_${TAG}_${N}gpu.json 2> gpurun_out/bench_${TAG}_${N}gpu.err
  echo "N=$N rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_${TAG}_${N}gpu.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])" || tail -5 gpurun_out/bench_${TAG}_${N}gpu.err
done

#!/bin/bash
# First GPU call of the next round: what the executor's knobs are worth on the box of the day (one GPU, ~25 s per point).
#   chunks   executor batches per 512-slot step (fill/drain of the pipeline vs per-batch gaps)
#   depth    batches in flight
#   back-sms back-end partition (the decode kernel got 1.6x faster in round 1e: re-balance)
TAG=${1:-r2a}
mkdir -p gpurun_out
run() {  # name, bench arguments
  timeout 200 python bench.py --cpu-slots 2 --e2e-slots 2 "${@:2}" > gpurun_out/sweep_${TAG}_$1.json 2>/dev/null
  python - "$1" "gpurun_out/sweep_${TAG}_$1.json" <<'P'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
s = d["roofline"]["stage_ms_per_launch"]
print("%-22s %8.0f slots/s  %.3f ms/step  k1 %.3f  back %.3f  part %s" % (sys.argv[1], d["value"], d["ms_per_step"], s["block_sums"],
      sum(v for k, v in s.items() if k != "block_sums"), d.get("run", d["config"]).get("sm_partition")))
P
}
for B in 24 32 40; do run back$B --back-sms $B; done
for C in 2 8; do run chunks$C --chunks $C; done
for D in 2 4; do run depth$D --depth $D; done
run chunks8_depth4 --chunks 8 --depth 4

#!/bin/bash
# A/B of the executor's back-end chain (run under gpurun): bench headline at forced back partitions with and without it.  usage: tools/gpu_chain_ab.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "pipe" 2>&1 | tail -2
run() {  # name chain args...
  name=$1; chain=$2; shift 2
  [ "$chain" = 1 ] && set -- --chain-back "$@"
  timeout 300 python bench.py --steps 40 --warmup 5 --no-configs --cpu-slots 4 "$@" > gpurun_out/bench_${TAG}_${name}.json 2> gpurun_out/bench_${TAG}_${name}.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}_${name}.json"))
st = d["roofline"]["stage_ms_per_launch"]
print("${name}: %.0f slots/s  %.4f ms/step  frac %.3f  part %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], {k: v for k, v in d["run"]["sm_partition"].items() if k != "chosen_by"}), {k: round(v, 3) for k, v in st.items()})
PY
}
run chain_28 1 --back-sms 28
run nochain_28 0 --back-sms 28
run nochain_26 0 --back-sms 26
run nochain_30 0 --back-sms 30
run auto 0

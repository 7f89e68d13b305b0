#!/bin/bash
# Quick GPU check (run under gpurun): parity suite, then one bench line; prints the stage times.  usage: tools/gpu_quick.sh TAG [bench args]
TAG=${1:-x}; shift
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12)
timeout 400 python bench.py --cpu-slots 8 "$@" > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}.json"))
print("value %.0f slots/s  %.4f ms/step  e2e %.1f  ok %d  launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["decoded_ok_slots_in_first_batch"], d["gpu_launches"]))
st = d["roofline"]["stage_ms_per_launch"]; print({k: round(v, 4) for k, v in st.items()}, "sum %.4f" % sum(st.values()), "frac %.3f" % d["roofline"]["frac"])
PY
tail -3 gpurun_out/bench_${TAG}.err

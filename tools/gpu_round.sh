#!/bin/bash
# Round-end style pass (run under gpurun): parity suite, both bench arms, then the ncu launch list of the bench command.
# usage: tools/gpu_round.sh TAG   -> gpurun_out/{pytest,bench,bench_ref,launches}_TAG.*
TAG=${1:-x}
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) | tee gpurun_out/pytest_${TAG}.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -3 gpurun_out/bench_${TAG}.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}_bench.csv \
    python bench.py --steps 2 --warmup 1 --cpu-slots 2 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}.json"))
print("value %.0f slots/s  %.4f ms/step  e2e %.1f  ok %s  launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("decoded_ok_slots_in_first_batch"), d["gpu_launches"]))
print(json.dumps(d["roofline"]))
print(open("gpurun_out/bench_ref_${TAG}.json").read()[:300])
PY

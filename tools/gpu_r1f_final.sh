#!/bin/bash
# round 1f evidence pass with the final kernels: parity, smoke, bench (partitioned + serial), the serial bench's ncu launch list,
# ncu --set full of the persistent node-centred decode kernel, the other BASELINE configurations, decode A/B, reference arm.
TAG=${1:-r1f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_${TAG}.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
timeout 600 python bench.py --back-sms 0 > gpurun_out/bench_${TAG}_serial.json 2> gpurun_out/bench_${TAG}_serial.err; echo "serial rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}_bench.csv \
    python bench.py --back-sms 0 --steps 2 --warmup 1 --cpu-slots 2 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1; echo "launch list rc=$?"
python tools/launch_summary.py gpurun_out/launches_${TAG}_bench.csv > gpurun_out/launches_${TAG}_bench.md 2>&1
export PROF_REPS=2 PROF_SLOTS=128
ncu --set full --clock-control none --import-source on -k regex:^decode_kernel -s 1 -c 1 -f -o gpurun_out/ncu_decode_kernel_${TAG} \
    python tools/prof_run.py > gpurun_out/ncu_decode_kernel_${TAG}.log 2>&1
ncu -i gpurun_out/ncu_decode_kernel_${TAG}.ncu-rep --page raw --csv > gpurun_out/ncu_raw_decode_kernel_${TAG}.csv 2>/dev/null
timeout 600 python tools/perf_configs.py ${TAG} > gpurun_out/perf_configs_${TAG}.log 2>&1; echo "configs rc=$?"
timeout 300 python tools/perf_decode_ab.py ${TAG} > gpurun_out/perf_decode_ab_${TAG}.log 2>&1; echo "ab rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; echo "ref rc=$?"
python - <<'P'
import json
for f in ("bench_r1f", "bench_r1f_serial", "bench_ref_r1f"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"], 1), d.get("ms_per_step"), d.get("e2e", {}).get("value"), (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"))
P
tail -4 gpurun_out/perf_configs_${TAG}.log | cut -c1-330

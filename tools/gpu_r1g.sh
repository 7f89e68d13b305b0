#!/bin/bash
# short evidence pass after the last decode change: bench (partitioned + serial), launch list, decode capture, decode A/B, audio path
TAG=${1:-r1g}
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
timeout 300 python bench.py --back-sms 0 > gpurun_out/bench_${TAG}_serial.json 2> gpurun_out/bench_${TAG}_serial.err; echo "serial rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}_bench.csv \
    python bench.py --back-sms 0 --steps 2 --warmup 1 --cpu-slots 2 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1; echo "launch list rc=$?"
python tools/launch_summary.py gpurun_out/launches_${TAG}_bench.csv > gpurun_out/launches_${TAG}_bench.md 2>&1
export PROF_REPS=2 PROF_SLOTS=128
ncu --set full --clock-control none --import-source on -k regex:^decode_kernel -s 1 -c 1 -f -o gpurun_out/ncu_decode_kernel_${TAG} \
    python tools/prof_run.py > gpurun_out/ncu_decode_kernel_${TAG}.log 2>&1
ncu -i gpurun_out/ncu_decode_kernel_${TAG}.ncu-rep --page raw --csv > gpurun_out/ncu_raw_decode_kernel_${TAG}.csv 2>/dev/null
timeout 300 python tools/perf_decode_ab.py ${TAG} > gpurun_out/perf_decode_ab_${TAG}.log 2>&1; echo "ab rc=$?"
timeout 200 python tools/perf_audio_ab.py > gpurun_out/perf_audio_ab_${TAG}.log 2>&1; tail -2 gpurun_out/perf_audio_ab_${TAG}.log
python - <<'P'
import json
for f in ("bench_r1g", "bench_r1g_serial"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"], 1), d.get("ms_per_step"), d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["stage_ms_per_launch"])
d = json.load(open("gpurun_out/perf_decode_ab_r1g.json"))
for k, v in d.items():
    print(k, {m: (round(v[m]["slots_per_s"]), min(v[m]["decode_ms"])) for m in ("nodes", "edges")}, v["decode_speedup"])
P
grep decode_kernel gpurun_out/launches_${TAG}_bench.md

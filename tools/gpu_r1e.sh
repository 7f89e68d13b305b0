#!/bin/bash
# round 1e GPU pass: parity (both BP kernels, C host program), decode A/B, bench (partitioned default + narrower back end)
TAG=${1:-r1e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_${TAG}.log
timeout 300 python tools/perf_decode_ab.py ${TAG} > gpurun_out/perf_decode_ab_${TAG}.log 2>&1; echo "ab rc=$?"; tail -4 gpurun_out/perf_decode_ab_${TAG}.log | cut -c1-1500
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['stage_ms_per_launch'], d['roofline']['frac'], d.get('run', d['config']).get('sm_partition'))"
timeout 300 python bench.py --back-sms 32 --cpu-slots 4 > gpurun_out/bench_${TAG}_back32.json 2> gpurun_out/bench_${TAG}_back32.err; echo "bench32 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}_back32.json')); print(d['value'], d['ms_per_step'], d['roofline']['stage_ms_per_launch'], d['roofline']['frac'], d.get('run', d['config']).get('sm_partition'))"

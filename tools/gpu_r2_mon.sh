#!/bin/bash
# 12 kHz monitor path check (run under gpurun, ONE GPU): its tests, then the bench line (roofline_extra carries the kernel's live time)
set -u
TAG=${1:-r2s}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "monitor or ft4 or real_recordings or wav or relinked or audio or cluster" > gpurun_out/pytest_mon_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_mon_${TAG}.log
bash tools/gpu_r2_bench.sh ${TAG}
python - <<PY
import json
b=json.load(open('gpurun_out/bench_${TAG}.json'))
for r in b['roofline_extra']:
    print(r['kernel'][:60], r.get('launch_ms'), r.get('us_per_slot'))
PY

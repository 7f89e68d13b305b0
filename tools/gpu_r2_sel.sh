#!/bin/bash
# Round-2 check (run under gpurun, ONE GPU): GPU tests, stage times, bench line.
set -u
TAG=${1:-r2j}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_${TAG}.log
timeout 300 python tools/perf_kernels.py ${TAG} > gpurun_out/perf_kernels_${TAG}.log 2>&1; echo "perf rc=$?"; tail -12 gpurun_out/perf_kernels_${TAG}.log
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_${TAG}.json

#!/bin/bash
# sync stage check (run under gpurun, ONE GPU): the tests that touch find_sync, stage times, launch lists.  usage: tools/gpu_r2_sync.sh TAG
set -u
TAG=${1:-r2n}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "selection or survives or min_score or sync_decode or golden or ft4 or monitor_dropin or dropin_find or real_recordings" > gpurun_out/pytest_sync_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_sync_${TAG}.log
timeout 300 python tools/perf_kernels.py ${TAG} > gpurun_out/perf_kernels_${TAG}.log 2>&1; echo "perf rc=$?"; tail -8 gpurun_out/perf_kernels_${TAG}.log | cut -c1-330
bash tools/gpu_r2_selprof.sh ${TAG} | grep -E "rc=|sync_" | awk '{print $1, $NF}' | sort | uniq -c

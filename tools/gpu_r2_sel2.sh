#!/bin/bash
# selection tests + stage times + launch lists of find_sync (run under gpurun, ONE GPU)
set -u
TAG=${1:-r2k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "selection or survives or min_score or sync_decode or golden or ft4 or monitor_dropin" > gpurun_out/pytest_sel_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_sel_${TAG}.log
timeout 300 python tools/perf_kernels.py ${TAG} > gpurun_out/perf_kernels_${TAG}.log 2>&1; echo "perf rc=$?"; tail -8 gpurun_out/perf_kernels_${TAG}.log
bash tools/gpu_r2_selprof.sh ${TAG} | grep -E "rc=|select" | awk '{print $1, $NF}' | sort | uniq -c
bash tools/gpu_r2_selsrc.sh | head -30

#!/bin/bash
# Source-level ncu captures (run under gpurun): one `--set full --import-source on` capture per kernel named on the command line,
# exported as raw + source CSV pages into gpurun_out/.  usage: tools/gpu_profile_src.sh TAG kernel [kernel ...]
set -u
TAG=$1; shift
mkdir -p gpurun_out
export PROF_SLOTS=${PROF_SLOTS:-128} PROF_REPS=2
for K in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:^${K} -s 1 -c 1 -f -o gpurun_out/ncu_${K}_${TAG} \
      python tools/prof_run.py > gpurun_out/ncu_${K}_${TAG}.log 2>&1
  ncu -i gpurun_out/ncu_${K}_${TAG}.ncu-rep --page raw --csv > gpurun_out/ncu_raw_${K}_${TAG}.csv 2>/dev/null
  ncu -i gpurun_out/ncu_${K}_${TAG}.ncu-rep --page source --csv > gpurun_out/ncu_src_${K}_${TAG}.csv 2>/dev/null
done
ls -la gpurun_out

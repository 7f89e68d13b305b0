#!/bin/bash
# launch lists of find_sync alone (tail / split selection) at 128 and 4096 slots (run under gpurun, ONE GPU)
set -u
TAG=${1:-r2j}
mkdir -p gpurun_out
for N in 128 4096; do
  PROF_SLOTS=$N PROF_NOISE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}_sync_${N}.csv \
      python tools/prof_sync.py > gpurun_out/prof_sync_${TAG}_${N}.log 2>&1; echo "N=$N rc=$?"
  grep -E "sync_|Memset" gpurun_out/launches_${TAG}_sync_${N}.csv | awk -F'","' '{print $5, $NF}' | tr -d '"' | tail -30
done

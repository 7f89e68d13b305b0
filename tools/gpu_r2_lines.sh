#!/bin/bash
# per-source-line profile of one kernel (run under gpurun, ONE GPU).  usage: tools/gpu_r2_lines.sh KERNEL TAG [top_n] [driver script: tools/prof_sync.py]
set -u
K=$1; TAG=$2; TOP=${3:-45}; DRV=${4:-tools/prof_sync.py}
export PROF_SLOTS=${PROF_SLOTS:-128} PROF_REPS=2
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:^${K} -s 1 -c 1 -f -o gpurun_out/ncu_${K}_${TAG} python $DRV > gpurun_out/ncu_${K}_${TAG}.log 2>&1
python tools/ncu_lines.py gpurun_out/ncu_${K}_${TAG}.ncu-rep $TOP > gpurun_out/ncu_lines_${K}_${TAG}.txt 2>&1
ncu -i gpurun_out/ncu_${K}_${TAG}.ncu-rep --page raw --csv > gpurun_out/ncu_raw_${K}_${TAG}.csv 2>/dev/null
rm -f gpurun_out/ncu_${K}_${TAG}.ncu-rep
cat gpurun_out/ncu_lines_${K}_${TAG}.txt

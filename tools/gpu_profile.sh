#!/bin/bash
# Round profile pass (run under gpurun): launch list of the whole path + `ncu --set full` captures of each kernel.
# Numbers printed under ncu are never bench values.  Outputs land in gpurun_out/ (copied to profiles/ by hand).
set -u
TAG=${1:-r1b}
mkdir -p gpurun_out
export PROF_SLOTS=${PROF_SLOTS:-32} PROF_REPS=2
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_${PROF_SLOTS}slots.csv \
    python tools/prof_run.py > gpurun_out/prof_run_${TAG}.log 2>&1
for K in cic_block_sums_kernel cic_comb_fir_kernel waterfall1024_kernel sync_score_ft8_kernel sync_select_kernel decode_kernel spots_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:^${K} -s 1 -c 1 -f -o gpurun_out/ncu_${K}_${TAG} \
      python tools/prof_run.py > gpurun_out/ncu_${K}_${TAG}.log 2>&1
  ncu -i gpurun_out/ncu_${K}_${TAG}.ncu-rep --page raw --csv > gpurun_out/ncu_raw_${K}_${TAG}.csv 2>/dev/null
done
ls -la gpurun_out

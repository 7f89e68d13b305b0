#!/bin/bash
# Launch list of the bench command WITHOUT the SM partition (Nsight Compute cannot profile kernels launched into CUDA green
# contexts: "Failed to prepare kernel for profiling"), the matching un-profiled bench line for share comparison, and the one
# kernel capture gpu_profile.sh missed.   usage: tools/gpu_round2.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
timeout 600 python bench.py --back-sms 0 --cpu-slots 8 > gpurun_out/bench_${TAG}_serial.json 2> gpurun_out/bench_${TAG}_serial.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}_bench.csv \
    python bench.py --back-sms 0 --steps 2 --warmup 1 --cpu-slots 2 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
export PROF_SLOTS=32 PROF_REPS=2
for K in sync_score_ft8_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:^${K} -s 1 -c 1 -f -o gpurun_out/ncu_${K}_${TAG} \
      python tools/prof_run.py > gpurun_out/ncu_${K}_${TAG}.log 2>&1
  ncu -i gpurun_out/ncu_${K}_${TAG}.ncu-rep --page raw --csv > gpurun_out/ncu_raw_${K}_${TAG}.csv 2>/dev/null
done
python tools/launch_summary.py gpurun_out/launches_${TAG}_bench.csv
python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}_serial.json')); print(d['value'], d['ms_per_step'], d['roofline']['stage_ms_per_launch'], d['roofline']['frac'])"

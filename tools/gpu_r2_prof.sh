#!/bin/bash
# Round-2 profile pass (run under gpurun, ONE GPU): launch list of the bench command (serial executor: Nsight Compute cannot attach to
# kernels launched into green contexts), `ncu --set full` captures of every kernel of both waterfall paths at 128 slots, raw pages,
# the summary JSON and the per-kernel instruction counts bench.py's roofline_extra uses.  Numbers printed under ncu are never bench values.
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
# steady state apart from the pipeline's fill and drain: 200 steps (800 batches) in one timed region
timeout 600 python bench.py --steps 200 --warmup 3 --no-configs --cpu-slots 1 > gpurun_out/bench_${TAG}_steps200.json 2> gpurun_out/bench_${TAG}_steps200.err; echo "steps200 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}_bench.csv \
    python bench.py --back-sms 0 --no-configs --steps 2 --warmup 3 --cpu-slots 1 --e2e-slots 2 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1; echo "launch list rc=$?"
python tools/launch_summary.py gpurun_out/launches_${TAG}_bench.csv > gpurun_out/launches_${TAG}_bench.md 2>&1
export PROF_REPS=2 PROF_SLOTS=128
for K in cic_block_sums_kernel cic_comb_fir_kernel waterfall1024_kernel sync_score_ft8_kernel sync_select_kernel decode_kernel spots_kernel synth_raw_kernel; do
  SKIP=1; [ $K = synth_raw_kernel ] && SKIP=0   # the input is synthesised once
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^${K} -s $SKIP -c 1 -f -o gpurun_out/ncu_${K}_${TAG} \
      python tools/prof_run.py > gpurun_out/ncu_${K}_${TAG}.log 2>&1
  ncu -i gpurun_out/ncu_${K}_${TAG}.ncu-rep --page raw --csv > gpurun_out/ncu_raw_${K}_${TAG}.csv 2>/dev/null
done
# the 12 kHz monitor path (ft8_lib's decode_ft8): its waterfall kernel, and the sync kernels at its geometry (960 bins, 137 232 positions)
export PROF_SLOTS=128 PROF_SIGNALS=20
for K in monitor_frames_kernel sync_score_ft8_kernel sync_select_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^${K} -s 1 -c 1 -f -o gpurun_out/ncu_${K}_12k_${TAG} \
      python tools/prof_audio.py > gpurun_out/ncu_${K}_12k_${TAG}.log 2>&1
  ncu -i gpurun_out/ncu_${K}_12k_${TAG}.ncu-rep --page raw --csv > gpurun_out/ncu_raw_${K}_12k_${TAG}.csv 2>/dev/null
done
# FT4: the generic per-plane scoring kernel (4 Costas arrays, 4 tones)
PROF_PROTOCOL=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:^sync_score_kernel -s 1 -c 1 -f -o gpurun_out/ncu_sync_score_kernel_ft4_${TAG} \
    python tools/prof_audio.py > gpurun_out/ncu_sync_score_kernel_ft4_${TAG}.log 2>&1
ncu -i gpurun_out/ncu_sync_score_kernel_ft4_${TAG}.ncu-rep --page raw --csv > gpurun_out/ncu_raw_sync_score_kernel_ft4_${TAG}.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/ncu_raw_*_${TAG}.csv --json gpurun_out/ncu_summary_${TAG}.json > gpurun_out/ncu_summary_${TAG}.txt 2>&1
python tools/ncu_inst.py gpurun_out/ncu_summary_${TAG}.json 128 gpurun_out/ncu_inst_${TAG}.json
rm -f gpurun_out/*.ncu-rep   # the raw pages and summaries are what is kept (the reports are tens of MB each)
ls -la gpurun_out | tail -40

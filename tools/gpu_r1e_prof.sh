#!/bin/bash
# round 1e evidence pass: serial bench line + its ncu launch list, ncu --set full of the node-centred decode kernel (and the
# edge-centred one for comparison), the other BASELINE configurations, the reference arm on the same box.
TAG=${1:-r1e}
mkdir -p gpurun_out
timeout 600 python bench.py --back-sms 0 --cpu-slots 8 > gpurun_out/bench_${TAG}_serial.json 2> gpurun_out/bench_${TAG}_serial.err; echo "serial rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}_bench.csv \
    python bench.py --back-sms 0 --steps 2 --warmup 1 --cpu-slots 2 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1; echo "launch list rc=$?"
python tools/launch_summary.py gpurun_out/launches_${TAG}_bench.csv > gpurun_out/launches_${TAG}_bench.md 2>&1; tail -12 gpurun_out/launches_${TAG}_bench.md
export PROF_REPS=2
for S in 32 128; do
  export PROF_SLOTS=$S
  ncu --set full --clock-control none --import-source on -k regex:^decode_kernel -s 1 -c 1 -f -o gpurun_out/ncu_decode_kernel_${TAG}_${S} \
      python tools/prof_run.py > gpurun_out/ncu_decode_kernel_${TAG}_${S}.log 2>&1
  ncu -i gpurun_out/ncu_decode_kernel_${TAG}_${S}.ncu-rep --page raw --csv > gpurun_out/ncu_raw_decode_kernel_${TAG}_${S}slots.csv 2>/dev/null
done
FT8B200_DECODE_VARIANT=1 ncu --set full --clock-control none -k regex:^decode_edges_kernel -s 1 -c 1 -f -o gpurun_out/ncu_decode_edges_kernel_${TAG}_128 \
    python tools/prof_run.py > gpurun_out/ncu_decode_edges_kernel_${TAG}.log 2>&1
ncu -i gpurun_out/ncu_decode_edges_kernel_${TAG}_128.ncu-rep --page raw --csv > gpurun_out/ncu_raw_decode_edges_kernel_${TAG}_128slots.csv 2>/dev/null
timeout 600 python tools/perf_configs.py ${TAG} > gpurun_out/perf_configs_${TAG}.log 2>&1; echo "configs rc=$?"; tail -5 gpurun_out/perf_configs_${TAG}.log | cut -c1-400
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_${TAG}.json
rm -f gpurun_out/*.ncu-rep.tmp
ls -la gpurun_out | tail -30

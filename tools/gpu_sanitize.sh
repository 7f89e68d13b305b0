#!/bin/bash
# compute-sanitizer passes over short whole-path runs (run under gpurun): memcheck, racecheck (shared-memory hazards of the
# FFT / sync / decode kernels), initcheck and synccheck.  Small batches: the tools serialise every kernel.  usage: tools/gpu_sanitize.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
: > gpurun_out/sanitizer_${TAG}.txt
export PROF_SLOTS=2 PROF_REPS=1 PROF_SIGNALS=20
SAN="compute-sanitizer --error-exitcode 9 --print-limit 20"
run() {  # name tool command...
  name=$1; tool=$2; shift 2
  timeout 600 $SAN --tool $tool "$@" > gpurun_out/san_${TAG}_${name}_${tool}.log 2>&1
  rc=$?
  echo "$name $tool rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_${TAG}_${name}_${tool}.log | tail -1)" | tee -a gpurun_out/sanitizer_${TAG}.txt
}
for tool in memcheck racecheck initcheck synccheck; do
  run raw $tool python tools/prof_run.py               # raw IQ -> decimator -> waterfall -> sync -> decode -> spots (bench input)
  run audio $tool python tools/prof_audio.py           # 12 kHz monitor path, FT8
done
PROF_PROTOCOL=0 run audio_ft4 memcheck python tools/prof_audio.py
PROF_PROTOCOL=0 run audio_ft4 racecheck python tools/prof_audio.py
run smoke memcheck python __graft_entry__.py smoke
run smoke racecheck python __graft_entry__.py smoke

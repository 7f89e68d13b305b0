#!/bin/bash
# one default bench run + a summary of its sub-measurements (run under gpurun, ONE GPU).  usage: tools/gpu_r2_bench.sh TAG
set -u
TAG=${1:-r2p}
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
b=json.load(open('gpurun_out/bench_${TAG}.json'))
print('value', b['value'], 'frac', b['roofline']['frac'], 'e2e', b['e2e']['value'], 'e2e_slots', b['e2e_slots']['value'])
for k,v in b['configs'].items():
    print(k, {kk:vv for kk,vv in v.items() if kk in ('slots_per_s','ms','ms_per_4096','parity','decoded_slots','subsystem_ms','receive_ms','wav_ms','wav_deferred_ms')})
PY

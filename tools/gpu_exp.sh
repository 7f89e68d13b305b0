#!/bin/bash
# Executor experiments (run under gpurun): one bench line per configuration, value + stage times printed.
mkdir -p gpurun_out
run() {  # name, env..., -- bench args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --cpu-slots 2 --e2e-slots 2 --steps 10 "$@" > gpurun_out/exp_${name}.json 2> gpurun_out/exp_${name}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/exp_${name}.json"))
    st = d["roofline"]["stage_ms_per_launch"]
    print("%-28s %8.0f slots/s  %.4f ms/step  k1 %.4f  back %.4f  ok %s" % ("${name}", d["value"], d["ms_per_step"], st["block_sums"], sum(v for k, v in st.items() if k != "block_sums"), d.get("decoded_ok_slots_in_first_batch")))
except Exception as e:
    print("${name} FAILED", e, open("gpurun_out/exp_${name}.err").read()[-300:])
PY
}
run init_front FT8B200_INIT_ON_FRONT=1 --
run default X=1 --
run free_front FT8B200_FREE_FRONT=1 --
run depth4 X=1 -- --depth 4
run free_front_depth4 FT8B200_FREE_FRONT=1 -- --depth 4
run back36 X=1 -- --back-sms 36
run back32 X=1 -- --back-sms 32
run chunks8 X=1 -- --chunks 8 --depth 4
run chunks2 X=1 -- --chunks 2

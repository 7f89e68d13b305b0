for S in 256 512; do
timeout 400 python bench.py --cpu-slots 4 --slots $S > gpurun_out/bench_s$S.json 2> gpurun_out/bench_s$S.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_s$S.json"))
print("slots $S value %.0f slots/s  %.4f ms/step  e2e %.1f  ok %d  launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["decoded_ok_slots_in_first_batch"], d["gpu_launches"]))
st = d["roofline"]["stage_ms_per_launch"]; print({k: round(v, 4) for k, v in st.items()}, "sum %.4f" % sum(st.values()), "frac %.3f" % d["roofline"]["frac"])
PY
tail -2 gpurun_out/bench_s$S.err
done

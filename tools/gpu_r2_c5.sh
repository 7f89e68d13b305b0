mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipe_streams" 2>&1 | tail -5
timeout 900 python bench.py > gpurun_out/bench_r2o.json 2> gpurun_out/bench_r2o.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2o.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_r2o.json'))
print(b['value'], b['roofline']['frac'], b['e2e']['value'])
print(b['configs']['c5_streams'])
PY

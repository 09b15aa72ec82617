timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_posterior.py -q -x 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 3 --cpu-budget 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench1 rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_c2.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e'].get('upload_seconds'), 'roofline', d['roofline']['frac'], d['roofline']['achieved'])
print(d['phases_ms'])
PY
tail -3 gpurun_out/bench_c2.err
timeout 1200 python tools/bench_configs.py c1 c4 k32 c3 > gpurun_out/bench_configs.jsonl 2> gpurun_out/bench_configs.err; echo "configs rc=$?"; cat gpurun_out/bench_configs.jsonl; tail -5 gpurun_out/bench_configs.err

timeout 900 python -m pytest tests/test_gpu_constrained.py -q -x 2>&1 | tail -25
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_c2.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['achieved'])
print(d['phases_ms'])
PY

timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_c2.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['achieved'])
print(d['phases_ms'])
PY
timeout 900 python tools/bench_configs.py c1 c4 k32 c3 2>&1 | grep -v "^$" | cut -c1-900 | tee gpurun_out/bench_configs.jsonl

timeout 1500 tools/gpu_isolated_tests.sh > /dev/null 2>&1; tail -1 gpurun_out/isolated_tests.log; grep FAIL gpurun_out/isolated_tests.log | head
timeout 1200 python bench.py --steps 20 --warmup 3 --cpu-budget 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_c2.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['achieved'])
print(d['phases_ms'])
PY
tail -5 gpurun_out/bench_c2.err

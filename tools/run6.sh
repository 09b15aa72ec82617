timeout 1500 tools/gpu_isolated_tests.sh > /dev/null 2>&1; tail -1 gpurun_out/isolated_tests.log; grep -A25 FAIL gpurun_out/isolated_tests.log | head -120
BTF_BAND_SCALAR=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "vs_oracle or golden" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 3 --cpu-budget 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench1 rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_c2.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['achieved'])
print(d['phases_ms'])
PY
tail -3 gpurun_out/bench_c2.err

timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_samplers.py tests/test_gpu_parity.py -q -x 2>&1 | tail -5
for mode in std alt; do
  if [ $mode = alt ]; then export BTF_STATS_K16_ALT=1; else unset BTF_STATS_K16_ALT; fi
  timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_$mode.json 2> gpurun_out/bench_c2_$mode.err; echo "bench $mode rc=$?"
  python - $mode <<'PY'
import json, sys
d = json.load(open('gpurun_out/bench_c2_%s.json' % sys.argv[1]))
print(sys.argv[1], 'value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['achieved'])
print(d['phases_ms'])
PY
done
unset BTF_STATS_K16_ALT
timeout 600 python tools/bench_configs.py c4 2>&1 | grep -v "^$" | cut -c1-700

timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x 2>&1 | tail -4
for mode in zpre nozpre; do
  if [ $mode = nozpre ]; then export BTF_STATS_NO_ZPRE=1; else unset BTF_STATS_NO_ZPRE; fi
  timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_$mode.json 2> gpurun_out/bench_c2_$mode.err; echo "bench $mode rc=$?"
  python - $mode <<'PY'
import json, sys
d = json.load(open('gpurun_out/bench_c2_%s.json' % sys.argv[1]))
print(sys.argv[1], 'value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['achieved'])
print(d['phases_ms'])
PY
done
unset BTF_STATS_NO_ZPRE
timeout 600 python tools/bench_configs.py k32 2>&1 | grep -v "^$" | cut -c1-700
timeout 1500 python bench.py --workload c5 --gpus 1 --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c5 slice value', d['value'], 'roofline', d['roofline']['frac'], d['roofline']['achieved']); print(d['phases_ms'])"

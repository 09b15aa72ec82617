timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 3 --cpu-budget 10 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_c2.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e'].get('host_buffers'), 'roofline', d['roofline']['frac'], d['roofline']['achieved'], 'cpu', d['cpu_baseline']['value'])
print(d['phases_ms'])
PY
timeout 900 python tools/bench_configs.py c1 c4 k32 c3 2>&1 | grep -v "^$" | cut -c1-900 | tee gpurun_out/bench_configs.jsonl
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_final2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:stats_kernel_zpre -s 3 -c 2 -o gpurun_out/prof_stats_zpre_r1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_stats.log 2>&1
echo "ncu stats rc=$?"
